"""Round-2 first GPU action: time the three host-buffer C-ABI calls (scripts/e2e_probe.py) under the opt-in host-path
knobs of capi.cu (all read once per process, hence one subprocess per setting).  Usage on the GPU box:
    python scripts/e2e_sweep.py > gpurun_out/e2e_sweep.txt
"""
import itertools
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SETTINGS = [{}, {"B200NB_HOST_TIMING": "1"}]     # the second run prints the per-phase breakdown to stderr (shown below)
for genes, workers in itertools.product(("6250", "12500", "25000"), ("2", "3")):
    SETTINGS.append({"B200NB_CHUNK_GENES": genes, "B200NB_CHUNK_WORKERS": workers})
SETTINGS += [{"B200NB_STAGE_CHUNK_MB": "4"}, {"B200NB_STAGE_CHUNK_MB": "8"}, {"B200NB_STAGE_CHUNK_MB": "32"},
             {"B200NB_DETECT_SF": "1"}, {"B200NB_DETECT_SF": "1", "B200NB_D2H_POPULATE": "1"},
             {"B200NB_D2H_HUGEPAGE": "1"}, {"B200NB_D2H_POPULATE": "1"},
             {"B200NB_D2H_POPULATE": "1", "B200NB_D2H_THREADS": "16"}, {"B200NB_D2H_THREADS": "16"}, {"B200NB_D2H_THREADS": "32"},
             {"B200NB_STAGE_THREADS": "16", "B200NB_D2H_THREADS": "32"},
             {"B200NB_D2H_HUGEPAGE": "1", "B200NB_D2H_THREADS": "16"},
             {"B200NB_CHUNK_GENES": "12500", "B200NB_CHUNK_WORKERS": "2", "B200NB_D2H_HUGEPAGE": "1"},
             {"B200NB_CHUNK_GENES": "12500", "B200NB_CHUNK_WORKERS": "3", "B200NB_D2H_HUGEPAGE": "1",
              "B200NB_D2H_THREADS": "16"}]
for s in SETTINGS:
    env = dict(os.environ, **s)
    r = subprocess.run([sys.executable, os.path.join(HERE, "e2e_probe.py")], env=env, capture_output=True, text=True,
                       timeout=600)
    line = (r.stdout.strip().splitlines() or [r.stderr.strip()[-300:]])[-1]
    if "B200NB_HOST_TIMING" in s:
        print("\n".join(l for l in r.stderr.splitlines() if l.startswith("b200nb timing"))[-1500:], flush=True)
    print(" ".join(f"{k[7:]}={v}" for k, v in s.items()) or "default", "|", line, flush=True)
