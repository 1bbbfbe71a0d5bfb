"""Learning-rate range test helpers used by `--test_lr` (reference: models/lr_finder.py): an exponential per-batch
schedule from the optimiser's lr up to `end_lr`, and the loss-vs-lr plot.  matplotlib is optional here: without it the
curve is written to lr_plot.csv instead of lr_plot.png."""
import numpy as np
from torch.optim.lr_scheduler import _LRScheduler


class BatchExponentialLR(_LRScheduler):
    """lr_i = base_lr * (end_lr / base_lr) ** (i / num_iter), stepped once per batch."""

    def __init__(self, optimizer, end_lr, num_iter, last_epoch=-1):
        self.end_lr, self.num_iter = end_lr, num_iter
        super().__init__(optimizer, last_epoch)

    def get_lr(self):
        frac = (self.last_epoch + 1) / self.num_iter
        return [b * (self.end_lr / b) ** frac for b in self.base_lrs]


def plot_lr(history, skip_start=10, skip_end=5, log_lr=True, show_lr=None):
    """Loss against learning rate, trimmed by `skip_start` / `skip_end` batches; marks the steepest descent."""
    if skip_start < 0 or skip_end < 0:
        raise ValueError("skip_start / skip_end cannot be negative")
    if show_lr is not None and not isinstance(show_lr, float):
        raise ValueError("show_lr must be float")
    stop = len(history["lr"]) - skip_end
    lrs, losses = history["lr"][skip_start:stop], history["loss"][skip_start:stop]
    steepest = None
    if len(losses) > 1:
        steepest = int(np.gradient(np.asarray(losses, dtype=np.float64)).argmin())
        print("Min numerical gradient: {:.2E}".format(lrs[steepest]))
    else:
        print("Failed to compute the gradients, there might not be enough points.")
    try:
        from matplotlib import pyplot as plt
        if not hasattr(plt, "plot"):
            raise ImportError("matplotlib stub")
    except ImportError:
        np.savetxt("lr_plot.csv", np.column_stack([lrs, losses]), delimiter=",", header="lr,loss", comments="")
        return
    plt.plot(lrs, losses)
    if log_lr:
        plt.xscale("log")
    plt.xlabel("Learning rate")
    plt.ylabel("Loss")
    if steepest is not None:
        plt.plot(lrs[steepest], losses[steepest], markersize=10, marker="o", color="red")
    if show_lr is not None:
        plt.axvline(x=show_lr, color="red")
    plt.savefig("lr_plot.png")
