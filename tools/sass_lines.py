"""Static SASS instruction counts per source line of one kernel (needs -lineinfo):
   cuobjdump -xelf all X.o; nvdisasm --print-line-info X.sm_100a.cubin > all.sass; sass_lines.py all.sass <substring of kernel name> [top]"""
import collections
import re
import sys

path, key = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
fn = None
line = None
cnt = collections.Counter()
ops = collections.defaultdict(collections.Counter)
for l in open(path):
    m = re.match(r'\s*\.text\.(\S+):', l)
    if m:
        fn = m.group(1)
        continue
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', l)
    if m:
        line = int(m.group(2))
        continue
    m = re.search(r'/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)', l)
    if m and fn and key in fn:
        cnt[line] += 1
        ops[line][m.group(2)] += 1
print("total", sum(cnt.values()))
for ln, n in sorted(cnt.items(), key=lambda x: -x[1])[:top]:
    print("%6d  L%-5d %s" % (n, ln, dict(ops[ln].most_common(6))))
