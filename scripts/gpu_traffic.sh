timeout 600 python -m pytest tests/test_gpu_parity.py -q -m "gpu and not slow" 2>&1 | tail -2
python bench.py --steps 500 --warmup 5 --cpu-seconds 0.2 | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'])"
ncu --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 8 -c 6 --csv --log-file gpurun_out/launches_warm.csv python bench.py --steps 4 --warmup 3 --cpu-seconds 0.1 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_warm.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
d={}
for r in rows[1:]: d.setdefault((int(r[ii]),r[ki][:28]),{})[r[mi].split('__')[1][:14]]=r[vi]
for k in sorted(d): print(k,d[k])
PY
