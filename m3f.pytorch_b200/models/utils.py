"""Metrics of the task module (reference: models/utils.py:6-26)."""


def concordance_cc2(r1, r2, reduction='mean'):
    """Batch sequence-wise CCC: biased covariance over unbiased variances (reference semantics)."""
    m1 = r1.mean(dim=-1, keepdim=True)
    m2 = r2.mean(dim=-1, keepdim=True)
    cov = ((r1 - m1) * (r2 - m2)).mean(dim=-1, keepdim=True)
    ccc = 2 * cov / (r1.var(dim=-1, keepdim=True) + r2.var(dim=-1, keepdim=True) + (m1 - m2) ** 2)
    if reduction == 'none':
        return ccc
    return ccc.mean()


def concordance_cc2_np(r1, r2):
    cov = ((r1 - r1.mean()) * (r2 - r2.mean())).mean()
    return 2 * cov / (r1.var() + r2.var() + (r1.mean() - r2.mean()) ** 2)


def mse(preds, labels):
    return sum((preds - labels) ** 2) / len(labels)
