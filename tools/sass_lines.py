#!/usr/bin/env python
"""Static SASS size of a kernel attributed to source lines (innermost line of a given file in the inline chain).
   python tools/sass_lines.py <obj.o> <kernel-substring> <source-file-basename> [top]"""
import collections, os, re, subprocess, sys, tempfile
obj, pat, fname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info-inline", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(dis) if l.startswith("//---") and ".text." in l and pat in l][0]
end = [i for i, l in enumerate(dis) if l.startswith("//---") and i > start][0]
cnt = collections.Counter(); ops = collections.defaultdict(collections.Counter); chain = []; new = True; paths = {}
for l in dis[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if new: chain, new = [], False
        chain.append((os.path.basename(m.group(1)), int(m.group(2)))); paths[os.path.basename(m.group(1))] = m.group(1); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?(\S+)", l)
    if m:
        new = True
        k = [c for c in chain if c[0] == fname]
        key = k[0] if k else ("(other)", 0)      # nvdisasm lists the innermost frame first
        cnt[key] += 1; ops[key][m.group(1).split(".")[0]] += 1
src = open(paths[fname]).read().split("\n") if fname in paths else []
print(sum(cnt.values()), "instructions in", pat)
for (f, ln), c in cnt.most_common(top):
    text = src[ln - 1].strip()[:100] if f == fname and ln <= len(src) else ""
    print(f"{c:5d} {f}:{ln:<4d} {' '.join(f'{o}:{n}' for o, n in ops[(f, ln)].most_common(4)):40s} | {text}")
