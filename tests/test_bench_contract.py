"""bench.py's reference arm (--impl reference: the per-sample CPU port on the host cores) runs without a
GPU, so its side of the JSON contract is checked here: one line, the same metric / unit / config keys
as the GPU arm, the e2e and cpu_baseline objects the tier asks for."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS=str(min(8, os.cpu_count() or 1)))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-samples", "2"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    j = json.loads(lines[0])
    assert j["impl"] == "reference"
    assert j["metric"] == "train samples/sec (fwd+bwd)" and j["unit"] == "samples/s" and j["higher_is_better"] is True
    assert j["n_gpus"] == 1 and j["steps"] == 1 and j["warmup"] == 0 and j["value"] > 0 and j["ms_per_step"] > 0
    assert j["vs_baseline"] is None and j["scaling"] == "weak" and j["data"] == "synthetic"
    assert "workload" in j["config"] and "configs[1]" in j["config"]["workload"] and "model" not in j["config"]
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and cb["sample"]
    e = j["e2e"]
    assert e["value"] == j["value"] and e["unit"] == j["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
