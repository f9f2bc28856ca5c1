#!/bin/bash
# Round-2 probe 2: instruction/barrier latencies, the rewritten enumerator (tiers, single staging copy), BKZ-60 with the
# enumeration trace and the finer LLL profile.
O=gpurun_out/r2
mkdir -p $O
echo "== latencies"; tests/_build/ubench_lat | tee $O/ubench_lat.txt
echo "== gpu tests"
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 > $O/t_all_2.log; tail -8 $O/t_all_2.log
echo "== BKZ-60 with enumeration trace"
rm -f /tmp/enum_trace.txt
B200_ENUM_TRACE=/tmp/enum_trace.txt timeout 300 python tools/gpurun_bkz60_trial.py > $O/bkz60_v2.txt 2>&1
grep "wall\|sec_\|status\|calls" $O/bkz60_v2.txt
python - <<'PY'
import re, collections
rows=[dict(t.split("=") for t in l.split()) for l in open("/tmp/enum_trace.txt")]
print("enum calls", len(rows))
tot=sum(float(r["total_us"]) for r in rows); host=sum(float(r["host_us"]) for r in rows); dev=sum(float(r["dev_ms"]) for r in rows)*1e3
print("total_us sum %.3f s, host breadth %.3f s, device %.3f s" % (tot/1e6, host/1e6, dev/1e6))
b=collections.defaultdict(lambda:[0,0.0,0.0,0.0,0])
for r in rows:
    n=int(r["dev_nodes"]); k=0 if n<1000 else 1 if n<10000 else 2 if n<100000 else 3 if n<1000000 else 4
    e=b[k]; e[0]+=1; e[1]+=float(r["total_us"]); e[2]+=float(r["host_us"]); e[3]+=float(r["dev_ms"])*1e3; e[4]+=n
for k in sorted(b): print("nodes<10^%d: calls %d mean total %.1f us host %.1f us dev %.1f us, nodes %d" % (k+3, b[k][0], b[k][1]/b[k][0], b[k][2]/b[k][0], b[k][3]/b[k][0], b[k][4]))
PY
cp /tmp/enum_trace.txt $O/enum_trace_bkz60.txt 2>/dev/null; gzip -f $O/enum_trace_bkz60.txt
echo "== BKZ-60 profile build"
B200_LIB_DIR=lib_prof timeout 400 python tools/gpurun_bkz60_trial.py > $O/bkz60_prof.txt 2>&1
grep -A8 "LLL profile" $O/bkz60_prof.txt; grep "wall\|sec_lll" $O/bkz60_prof.txt
echo "== bench (no bkz)"
timeout 300 python bench.py --no-bkz --no-cpu-baseline > $O/bench_v2.json 2> $O/bench_v2.err; python -c "
import json; j=json.loads(open('$O/bench_v2.json').read().strip().splitlines()[-1]); print(j['value'], j['roofline']['frac'], j['e2e']['value'], j.get('enum'))"
echo done
