"""dev helper: build a variant of the library next to the product one (kernel experiments; load it with SPI_B200_LIB=...):
    python tools/build_variant.py NAME [-DFLAG ...]   ->  tools/_build/libspi_b200_NAME.so"""
import subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from spi_active_b200 import _lib
out = ROOT / "tools" / "_build" / f"libspi_b200_{sys.argv[1]}.so"
out.parent.mkdir(exist_ok=True)
cmd = [_lib._nvcc(), *_lib.NVCC_FLAGS, *sys.argv[2:], "-Xptxas", "-v", "-o", str(out), *map(str, _lib.SOURCES)]
res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
lines = res.stdout.splitlines()
for i, l in enumerate(lines):
    if "Compiling entry function" in l and "rollout_ws" in l:
        print(l.split("'")[1], "|", lines[i + 2].strip() if i + 2 < len(lines) else "", "|", lines[i + 1].strip())
if res.returncode:
    print(res.stdout[-3000:]); sys.exit(1)
print("built", out)
