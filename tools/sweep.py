"""BASELINE configs[4]: num_envs sweep of the rollout step on this rank's GPU(s): one bench.py line per size.

    python tools/sweep.py [--gpus N] > profiles/r01_sweep_nN.jsonl
"""
import json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
gpus = int(sys.argv[sys.argv.index("--gpus") + 1]) if "--gpus" in sys.argv else 1
for n in (1024, 2048, 4096, 8192, 16384, 32768, 65536):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", str(gpus), "--envs-per-gpu", str(n), "--steps", str(max(40, 200 * 4096 // n)), "--warmup", "10", "--no-cpu-baseline", "--no-extra"]
    if gpus > 1:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(gpus), "--master-addr", "127.0.0.1", "--master-port", "29533"] + cmd[1:]
    r = subprocess.run(cmd, capture_output=True, text=True)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if not line:
        print(json.dumps({"envs_per_gpu": n, "error": r.stderr[-300:]}), flush=True)
        continue
    j = json.loads(line[-1])
    print(json.dumps({"envs_per_gpu": n, "n_gpus": gpus, "value": j["value"], "ms_per_step": j["ms_per_step"], "kernel_ms": j["roofline"]["kernel_ms"],
                      "e2e": j["e2e"]["value"], "hbm_frac": j["roofline"]["frac"], "fp32_frac": j["roofline"]["fp32_frac"]}), flush=True)
