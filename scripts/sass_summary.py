"""Static evidence from the built library (no GPU needed): per kernel, the SASS instruction mix that matters for
the design claims -- bulk (TMA-engine) row copies (UBLKCP), cp.async gathers (LDGSTS), mbarrier waits / arrives
(SYNCS), warp reductions (REDUX), shuffles (SHFL), CTA barriers (BAR), FP64 arithmetic (DFMA / DADD / DMUL), FP64
reciprocal seeds (MUFU.RCP64H), local-memory traffic (LDL / STL: spills).
    python scripts/sass_summary.py > profiles/r02g_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, 'cgpm_b200', 'libcgpm_b200.so')
sass = subprocess.run(['/usr/local/cuda/bin/cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
pats = collections.OrderedDict([
    ('UBLKCP', r'\bUBLKCP'), ('LDGSTS', r'\bLDGSTS'), ('SYNCS', r'\bSYNCS'), ('C?REDUX', r'\bC?REDUX'), ('SHFL', r'\bSHFL'),
    ('BAR', r'\bBAR\.'), ('DFMA', r'\bDFMA'), ('DADD', r'\bDADD'), ('DMUL', r'\bDMUL'), ('MUFU.RCP64H', r'MUFU\.RCP64H'),
    ('LDS', r'\bLDS'), ('LDG', r'\bLDG'), ('LDL', r'\bLDL'), ('STL', r'\bSTL')])
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None or '/*' not in line:
        continue
    if re.search(r'^\s+/\*[0-9a-f]{4,}\*/\s+\S', line):
        total[cur] += 1
        for k, p in pats.items():
            if re.search(p, line):
                counts[cur][k] += 1
demangle = subprocess.run(['c++filt'] + list(counts), capture_output=True, text=True).stdout.splitlines()
print('kernel | instructions | ' + ' | '.join(pats))
for name, nice in zip(counts, demangle):
    nice = re.sub(r'\(anonymous namespace\)::', '', nice)
    nice = re.sub(r'\(.*', '', nice)
    print('%s | %d | %s' % (nice, total[name], ' | '.join(str(counts[name][k]) for k in pats)))
