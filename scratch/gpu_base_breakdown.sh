for p in 3 30; do
PROBE_N=100 PROBE_P=$p ABM_BASE_FUSED=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 90 --csv --log-file gpurun_out/bb_$p.csv python scratch/base_n100_probe.py 100 > /dev/null 2>&1
python - <<PY
import csv, collections
rows=list(csv.reader(open("gpurun_out/bb_$p.csv")))
h=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
hdr=rows[h]; kn=hdr.index('Kernel Name'); mv=hdr.index('Metric Value'); mu=hdr.index('Metric Unit')
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[h+1:]:
    if len(r)!=len(hdr): continue
    v=float(r[mv].replace(',',''));u=r[mu]
    v = v/1e3 if u in ('ns','nsecond') else (v if u in ('us','usecond') else v*1e3)
    a=agg[r[kn][:40]]; a[0]+=1; a[1]+=v
for k,a in sorted(agg.items(), key=lambda x:-x[1][1])[:5]: print("P=$p", k, a[0], "avg %.1f us" % (a[1]/a[0]))
PY
done
