"""GPU parity tests proper (`pytest -m gpu`): every case of tests/gpu_cases.py, through the C ABI, against the oracle.

What the error keys mean and which tolerance applies (tests/gpu_cases.py TOL / TOLS):
  out / dw / sum / sumsq      raw tensor-core ops vs fp32 CPU evaluation of the same op on bf16-rounded operands
                              (bf16 output rounding 2^-8 -> 1.5e-2 range-normalised; fp32 outputs are ~1e-6)
  out_emu / grad_* / d_* / dx the CUDA path vs the oracle with bf16 *storage* emulation (same rounding points):
                              differences reduce to summation order (forward 1.5e-2, gradients 3e-2 relative L2)
  out_ref / va_ref / loss_ref the CUDA path vs the golden output of the UNMODIFIED reference module (fp32); the
                              north-star bf16 bar is 2e-2 on V/A predictions; `floor` reports what the bf16-emulating
                              oracle itself loses on the same (deliberately high-gain) synthetic weights
"""
import pytest
import torch

from tests import gpu_cases as G

pytestmark = pytest.mark.gpu

# Whole-network train-mode gradients on the tiny golden batches are numerically chaotic (rounding only the conv weights
# to bf16 in exact fp64 math already moves them by 30-45 %, DESIGN.md "Numerics"); their backward chain is covered by
# the well-conditioned block_* / convnd_* / golden_{gru,attfusion,tcn} cases instead, and only forward/loss parity is
# asserted for them here.
CHAOTIC_GRADS = {"golden_resnet_trunk_train", "golden_va3dresnet_train", "golden_av_resnet_attention_train"}


@pytest.mark.parametrize("name", sorted(G.CASES))
def test_case(name):
    assert torch.cuda.is_available()
    errs = G.run_case(name)
    if name in CHAOTIC_GRADS:
        errs = {k: v for k, v in errs.items() if not k.startswith("grad")}
    bad = {k: v for k, v in errs.items() if not isinstance(v, dict) and (v != v or v >= G.TOLS.get(k, G.TOL))}
    assert not bad, (name, bad, errs)


def test_library_loaded_and_counting():
    from m3t_b200 import lib, raw
    n0 = lib.launch_count()
    a = torch.randn(128, 64, device="cuda").bfloat16()
    raw.gemm(a, a)
    torch.cuda.synchronize()
    assert lib.launch_count() == n0 + 1


def test_ccc_three_decimals():
    """CCC of the CUDA predictions vs synthetic labels equals the reference's to 3 decimals (north-star criterion)."""
    from tests.golden_util import load, ref_batch
    from m3t_b200.models.utils import concordance_cc2
    fx = load("av_v2psplit_attention_eval")
    m = G._build(fx).eval()
    with torch.no_grad():
        out = m(ref_batch(fx["inputs"], "cuda")).float().cpu()
    for ch, lab in ((7, "label_valence"), (8, "label_arousal")):
        y = fx["inputs"][lab].reshape(-1)
        c_ref = float(concordance_cc2(fx["out"][..., ch].reshape(-1), y, "none"))
        c_gpu = float(concordance_cc2(out[..., ch].reshape(-1), y, "none"))
        assert abs(c_ref - c_gpu) < 1e-3, (ch, c_ref, c_gpu)
