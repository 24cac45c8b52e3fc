"""Shared test helpers: deterministic weights (key-name generator + calibrated BN statistics) and tolerances."""
import json
import os

import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
CFG_N = os.path.join(ROOT, "yolo-master_b200", "cfg", "models", "26", "yolo26-master-n.yaml")

# BASELINE.json north_star: fp16 conv/attn outputs within 1e-3 abs / 1e-2 rel of the reference forward
ATOL, RTOL = 1e-3, 1e-2


def yaml_n():
    return yaml.safe_load(open(CFG_N))


def yaml_of(rel):
    """A model YAML of this package's cfg tree, e.g. 'master/v0/det/yolo-master-n.yaml'."""
    return yaml.safe_load(open(os.path.join(ROOT, "yolo-master_b200", "cfg", "models", rel)))


def synth_sd_from_keys(seed=0, name="yolo26-master-n"):
    """fp32 CPU state_dict rebuilt from the reference key table only (no model code involved)."""
    from yolo_master_b200.utils.synth import fill_state_dict_, load_norm_stats_

    keys = json.load(open(os.path.join(GOLD, f"{name}.keys.json")))
    dt = {"torch.float32": torch.float32, "torch.int64": torch.int64}
    sd = {k: torch.zeros(shape, dtype=dt[d]) for k, (shape, d) in keys.items()}
    fill_state_dict_(sd, seed)
    load_norm_stats_(sd, torch.load(os.path.join(GOLD, f"{name}.bnstats.pt")))
    return sd


def close_stats(a, b, atol=ATOL, rtol=RTOL):
    """(max abs err, fraction of elements violating |a-b| <= atol + rtol*|b|)."""
    a, b = a.float().cpu(), b.float().cpu()
    err = (a - b).abs()
    bad = err > (atol + rtol * b.abs())
    return float(err.max()) if err.numel() else 0.0, float(bad.float().mean()) if err.numel() else 0.0


def assert_close(a, b, atol=ATOL, rtol=RTOL, max_bad_frac=0.0, what=""):
    mx, bad = close_stats(a, b, atol, rtol)
    assert bad <= max_bad_frac, f"{what}: max abs err {mx:.3e}, {bad * 100:.4f}% of elements outside atol={atol} rtol={rtol}"


def assert_within_noise(y, ref, sim, what="", k_mean=2.0, k_max=3.0, outlier_frac=0.0):
    """Chained-op criterion: the CUDA path's deviation from the fp32 oracle must stay within a small multiple of the
    deviation of the oracle's own fp16-storage model (`oracle.fp16_storage()`), i.e. within the noise ANY fp16 execution
    of the same graph has.  Single kernels are held to the strict north-star tolerance instead (assert_close)."""
    y, ref, sim = y.float().cpu(), ref.float(), sim.float()
    e, n = (y - ref).abs(), (sim - ref).abs()
    rms = float(ref.pow(2).mean().sqrt())
    assert float(e.mean()) <= k_mean * float(n.mean()) + 2e-4 * rms, \
        f"{what}: mean err {float(e.mean()):.3e} vs fp16-storage noise {float(n.mean()):.3e} (rms {rms:.3f})"
    emax = float(e.max())
    if outlier_frac > 0 and e.numel() > 1:
        # chains THROUGH a per-token discrete router: a token whose top-k margin is below the fp16 noise of its input may
        # legitimately take another expert (in the reference's own fp16 execution too), which moves that token by O(1).
        # The max criterion is then applied to the (1 - outlier_frac) quantile; the mean criterion above still covers all.
        emax = float(e.flatten().kthvalue(max(1, int(e.numel() * (1.0 - outlier_frac))))[0])
    assert emax <= k_max * float(n.max()) + 2e-3 * max(1.0, rms), \
        f"{what}: max err {emax:.3e} (outlier_frac={outlier_frac}) vs fp16-storage noise max {float(n.max()):.3e} (rms {rms:.3f})"
