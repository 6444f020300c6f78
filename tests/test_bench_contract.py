"""bench.py contract pieces that run without a GPU: the reference arm (the reference's own core, oracle/_ref, timed on
the host) prints ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "liblpmp_ref.so")), reason="oracle/_ref has not been built")
def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--crop", "48"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and len(lines) == 1, r.stderr[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "msg_updates/s"
    assert d["metric"].startswith("pairwise msg updates/s") and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("C3") and "sample" in d["config"] and d["scaling"] == "strong"
    # the arm exits normally (no os._exit) and keeps the reference's exit-time pool statistics off stdout
    assert len([l for l in r.stdout.splitlines() if l.strip()]) == 1


def test_bench_defaults_follow_the_contract():
    src = open(os.path.join(ROOT, "bench.py")).read()
    for needle in ('"--gpus", type=int, default=1', '"--steps"', '"--warmup"', '"--impl"', '"roofline"', '"cpu_baseline"', '"e2e"',
                   '"gpu_launches"', '"clocks"', '"time_to_bound"'):
        assert needle in src, needle
    assert "os._exit" not in src and 'default="C3"' in src and '"--scaling", default="strong"' in src
    assert "oracle" not in open(os.path.join(ROOT, "lp_mp_b200", "lp.py")).read().lower()
