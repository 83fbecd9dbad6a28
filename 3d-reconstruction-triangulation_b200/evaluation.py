"""Accuracy statistics of predicted paths against ground-truth positions: the
batched counterpart of the reference's offline evaluator (src/test.cpp:57-87, src/utils.cpp:133-213).
Runs in torch on whatever device the paths live on; predicted points of (0,0,0) mean "no point in that
frame" and are skipped exactly like the reference does (but, like the reference, still counted in the
denominators of mean and variance)."""
from __future__ import annotations


def path_errors(pred, label):
    """pred, label: [F,3] tensors -> (errors of the frames that have a point, F)."""
    import torch
    n = min(pred.shape[0], label.shape[0])
    p, l = pred[:n], label[:n]
    have = ~((p[:, 0] == 0) & (p[:, 1] == 0) & (p[:, 2] == 0))  # utils.cpp:138-139
    err = torch.sqrt(((p - l) ** 2).sum(dim=1))
    return err[have & ~torch.isnan(err)], n


def path_statistics(pred, label):
    """mean (utils.cpp:133-148), 'std' = variance over all frames (:150-167), median (:191-213),
    quartile deviation (:169-189) of one predicted path against one label path."""
    import torch
    err, n = path_errors(pred, label)
    if err.numel() == 0:
        raise RuntimeError("There are no errors to calculate median from")
    mean = float(err.sum() / n)
    var = float(((err - mean) ** 2).sum() / n)
    s, _ = torch.sort(err)
    m = s.numel()
    median = float(s[(m - 1) // 2]) if m % 2 == 1 else float((s[m // 2] + s[m // 2 - 1]) / 2)
    q1, q3 = s[min(n // 4, m - 1)], s[min(3 * (n // 4), m - 1)]
    return dict(mean=mean, std=var, median=median, quartile_deviation=float((q3 - q1) / 2), frames_with_point=int(m), frames=n)


def evaluate(pred_paths, label_paths):
    """For every predicted path the label path with the smallest mean error (src/test.cpp:66-84) and its
    statistics.  pred_paths [D,F,3], label_paths [L,F,3] in the same unit."""
    out = []
    for d in range(pred_paths.shape[0]):
        best = None
        for k in range(label_paths.shape[0]):
            try:
                st = path_statistics(pred_paths[d], label_paths[k])
            except RuntimeError:
                continue
            if best is None or st["mean"] < best["mean"]:
                best = dict(st, label=k)
        out.append(best)
    return out
