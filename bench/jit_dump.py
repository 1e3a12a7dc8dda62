"""Developer tool (no GPU): write the specialised Q1 scan kernel's source, cubin and SASS to a directory and print the
instruction mix of the per-tile loop.  python bench/jit_dump.py /tmp/jit [ngroups]"""
import collections
import pathlib
import re
import subprocess
import sys

root = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(root))
sys.path.insert(0, str(root / "tests"))
import test_jit  # noqa: E402

out = pathlib.Path(sys.argv[1] if len(sys.argv) > 1 else "/tmp/jit")
out.mkdir(parents=True, exist_ok=True)
ng = int(sys.argv[2]) if len(sys.argv) > 2 else 3
masked = (sys.argv[3] != "exact") if len(sys.argv) > 3 else True
src = test_jit.q1_source(out, ng, masked)
(out / "q1.cu").write_text(src)
cubin = test_jit.compile_source(src)
(out / "q1.cubin").write_bytes(cubin)
sass = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", str(out / "q1.cubin")], capture_output=True, text=True, check=True).stdout
(out / "q1.sass").write_text(sass)
res = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-res-usage", str(out / "q1.cubin")], capture_output=True, text=True).stdout
print(res.strip().splitlines()[-1] if res.strip() else "")
# instruction mix of the main function
body = sass.split("Function : msc_jit_dense")[1] if "Function : msc_jit_dense" in sass else sass
ops = re.findall(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", body)
mix = collections.Counter(o.split(".")[0] for o in ops)
print(len(ops), "SASS instructions in msc_jit_dense")
print(", ".join(f"{k} {v}" for k, v in mix.most_common(30)))
