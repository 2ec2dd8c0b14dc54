"""bench.py's control flow on the emulator (tests/bench_dry_run.py).  About 7 minutes per run, so it only runs when
LZF_RUN_SLOW=1; the builder ran both at the end of round 2."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
slow = pytest.mark.skipif(os.environ.get("LZF_RUN_SLOW") != "1", reason="slow (~7 min): set LZF_RUN_SLOW=1")


def _line(out):
    lines = [ln for ln in out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out[-2000:]
    return json.loads(lines[0])


def _no_errors(o, path=""):
    if isinstance(o, dict):
        for k, v in o.items():
            assert k not in ("error", "skipped"), (path, k, v)
            _no_errors(v, path + "/" + k)


@slow
def test_bench_control_flow_single_rank():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "bench_dry_run.py")], capture_output=True, text=True, timeout=3000)
    assert p.returncode == 0, p.stderr[-3000:]
    d = _line(p.stdout)
    _no_errors(d)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks", "compress", "extra_configs", "single_file"):
        assert key in d, key
    assert d["gpu_launches"] == d["steps"] and {"h2d_bytes_per_step", "d2h_bytes_per_step", "ceiling_gbs"} <= set(d["e2e"])


@slow
def test_bench_control_flow_two_ranks_over_gloo():
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", os.path.join(ROOT, "tests", "bench_dry_run.py"), "--gpus", "2"],
                       capture_output=True, text=True, timeout=3000)
    assert p.returncode == 0, p.stderr[-3000:]
    d = _line(p.stdout)
    _no_errors(d)
    assert d["n_gpus"] == 2
    g, x = d["compress"]["gather"], d["extra_configs"]["config4"]["exchange"]
    assert g["verified"] == {"archive_slices_equal_senders_digests": True, "scattered_frames_equal_on_every_rank": True}
    assert x["verified"]["archive_slices_equal_senders_digests"] and x["verified"]["scattered_frames_equal_on_every_rank_and_decode_bit_exact"]
