#!/bin/bash
# session 2, run K: launch list (serialised, cold cache) + DRAM bytes of every kernel of one bench step, default flags
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
CMD="python bench.py --steps 1 --warmup 1 --cpu-chunks 0 --batch 64 --no-stats"
$CMD > gpurun_out/plain_k.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_b64.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?" >> gpurun_out/summary.txt
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_b64.csv')) if len(r)>10]
hdr=rows[0]; idx={h:i for i,h in enumerate(hdr)}
agg=collections.OrderedDict()
for r in rows[1:]:
    name=r[idx['Kernel Name']].split('(')[0][-60:]; m=r[idx['Metric Name']]; v=float(r[idx['Metric Value']].replace(',',''))
    unit=r[idx['Metric Unit']]
    if m=='gpu__time_duration.sum':
        if unit=='ns': v/=1e3
        elif unit=='ms': v*=1e3
        elif unit=='s': v*=1e6
    else:
        mult={'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}.get(unit,1); v*=mult
    a=agg.setdefault(name,dict(n=0,us=0.0,rd=0.0,wr=0.0))
    if m=='gpu__time_duration.sum': a['n']+=1; a['us']+=v
    elif 'read' in m: a['rd']+=v
    else: a['wr']+=v
tot=sum(a['us'] for a in agg.values())
print('total ms %.1f'%(tot/1e3))
for k,a in sorted(agg.items(), key=lambda t:-t[1]['us']):
    print('%-62s n %4d  ms %8.2f  %5.1f%%  avg us %8.1f  dram rd GB %7.2f wr GB %7.2f'%(k,a['n'],a['us']/1e3,100*a['us']/tot,a['us']/a['n'],a['rd']/1e9,a['wr']/1e9))
PY
cat gpurun_out/summary.txt
