#!/usr/bin/env python
"""SASS size of a kernel attributed to source regions (needs -lineinfo).  Usage:
   python tools/sass_breakdown.py allocnet_b200/build/kernels_s3_l8.o optimize_kernel"""
import collections, os, re, subprocess, sys, tempfile
obj, pat = sys.argv[1], sys.argv[2]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info-inline", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(dis) if l.startswith("//---") and ".text." in l and pat in l][0]
end = [i for i, l in enumerate(dis) if l.startswith("//---") and i > start][0]
src = {}
def fn_of(f, ln):
    """enclosing function name by scanning the source upward for a definition line"""
    path = f
    if path not in src:
        try: src[path] = open(path).read().split("\n")
        except OSError: src[path] = []
    L = src[path]
    for i in range(min(ln, len(L)) - 1, -1, -1):
        m = re.match(r"^(?:__device__|__global__|template|static|inline|__host__).*?\b([A-Za-z_]\w*)\s*\(", L[i])
        if m and not L[i].startswith("template"):
            return m.group(1)
    return os.path.basename(f)
chain, new = [], True
tot = collections.Counter(); ops = collections.defaultdict(collections.Counter)
for l in dis[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if new: chain, new = [], False
        chain.append((m.group(1), int(m.group(2)))); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?(\S+)", l)
    if m:
        new = True
        names = [fn_of(f, ln) for f, ln in chain]
        # innermost function that is not a tiny helper
        small = {"sh_up", "sh_dn", "sh_xor", "group_sum", "group_max", "cfact", "cfall", "load_plane", "smoothed_l1_pos",
                 "inv_small", "forward_t", "backward_grad_t", "gdot", "ginf", "load_x", "store_x", "view_of", "fma", "min", "max"}
        name = next((n for n in names if n not in small), names[0] if names else "?")
        tot[name] += 1; ops[name][m.group(1).split(".")[0]] += 1
n = sum(tot.values())
print(f"{pat}: {n} instructions = {n*16/1024:.0f} KB")
for k, v in tot.most_common():
    print(f"  {k:24s} {v:6d}  " + " ".join(f"{o}:{c}" for o, c in ops[k].most_common(7)))
