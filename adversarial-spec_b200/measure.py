"""Turning the engine's in-situ kernel timeline into per-kernel costs.

`Engine.ktrace_read()` returns (timestamp_ns, kind) for block 0 of every decode-path kernel,
in launch order.  Kernels of a step are serialised by data dependencies, so the cost of kernel
i — its run time plus the launch gap behind it — is stamp[i+1] - stamp[i].  This is measured
inside the CUDA-graph replay, at production clocks and cache state.
"""

from __future__ import annotations

from collections import OrderedDict

KIND_NAMES = {1: "gemv", 2: "attn", 3: "sample_merge", 4: "rope", 5: "attn_combine", 6: "sample_scan"}
SAMPLE_MERGE = 3


def split_steps(trace: list[tuple[int, int]]) -> list[list[tuple[int, int]]]:
    """Steps start at the sampler's merge kernel (kind 3).  Returns, per complete step, the list of
    (cost_ns, kind) of its kernels; the first and last partial steps are dropped."""
    starts = [i for i, (_, k) in enumerate(trace) if k == SAMPLE_MERGE]
    steps = []
    for a, z in zip(starts[:-1], starts[1:]):
        seg = trace[a: z + 1]
        steps.append([(seg[i + 1][0] - seg[i][0], seg[i][1]) for i in range(len(seg) - 1)])
    return steps[1:] if len(steps) > 1 else steps


def summarize(trace: list[tuple[int, int]], n_layers: int) -> dict:
    steps = split_steps(trace)
    if not steps:
        return {"steps": 0}
    n = len(steps[0])
    steps = [s for s in steps if len(s) == n]
    per_layer = (n - 3) // n_layers if n > 3 else 0
    rows: "OrderedDict[str, list[float]]" = OrderedDict()
    for s in steps:
        for i, (dt, kind) in enumerate(s):
            if i == 0:
                key = "sample merge + next embedding"
            elif i == n - 1:
                key = "sample vocabulary scan"
            elif i == n - 2:
                key = "gemv lm_head"
            else:
                key = f"layer[{(i - 1) % per_layer}] {KIND_NAMES.get(kind, kind)}"
            rows.setdefault(key, []).append(float(dt))
    total = sum(sum(v) for v in rows.values()) / len(steps)
    out_rows = OrderedDict()
    gemv_ns = 0.0
    gemv_launches = 0
    for k, v in rows.items():
        per_step = sum(v) / len(steps)
        cnt = len(v) / len(steps)
        out_rows[k] = {"launches_per_step": cnt, "us_each": per_step / cnt / 1e3, "us_per_step": per_step / 1e3,
                       "share": per_step / total}
        if "gemv" in k:
            gemv_ns += per_step
            gemv_launches += int(round(cnt))
    return {"steps": len(steps), "kernels_per_step": n, "us_per_step": total / 1e3, "rows": out_rows,
            "gemv_us_per_step": gemv_ns / 1e3, "gemv_launches_per_step": gemv_launches}
