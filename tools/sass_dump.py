"""cuobjdump -sass of the built library: per-kernel opcode histograms (all kernels) and full listings of the hot ones.
usage: python tools/sass_dump.py <out_dir> [round-tag]"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pttools_b200", "libpttools_b200.so")
out_dir, tag = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "r02")
HOT = ("sdv_gemm_tma_kernel", "sdv_gemm_kernel", "spec_den_gw_cells_kernel", "sin_transform_moments_kernel", "a2_moments_kernel", "sdv_weights_kernel")
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
kernels, cur = collections.OrderedDict(), None
for ln in txt.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        kernels[cur] = []
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        kernels[cur].append(ln.rstrip())
with open(os.path.join(out_dir, f"{tag}_sass_opcodes.txt"), "w") as f:
    f.write(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a): instructions per kernel by opcode\n")
    for name, lines in kernels.items():
        ops = collections.Counter()
        for ln in lines:
            toks = [t for t in ln.split("*/", 1)[1].split() if not t.startswith("@")]
            if toks:
                ops[toks[0].rstrip(";")] += 1
        f.write(f"\n{name}: {len(lines)} instructions\n  " + "  ".join(f"{o} {c}" for o, c in ops.most_common(18)) + "\n")
for name, lines in kernels.items():
    short = name.split("::")[-1]
    if short in HOT:
        with open(os.path.join(out_dir, f"{tag}_sass_{short}.txt"), "w") as f:
            f.write(f"# cuobjdump -sass, {name} (sm_100a)\n" + "\n".join(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", ln) for ln in lines) + "\n")
print("kernels:", len(kernels))
