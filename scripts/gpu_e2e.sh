python - <<'PY'
import torch, time
x = torch.randn(64,1,160000).pin_memory(); d = torch.empty_like(x, device="cuda"); o = torch.empty(64,1,128,313, device="cuda"); oh = torch.empty(64,1,128,313).pin_memory()
for _ in range(3): d.copy_(x, non_blocking=True); oh.copy_(o, non_blocking=True)
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(20): d.copy_(x, non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/20
print("H2D 41MB: %.3f ms  %.1f GB/s" % (dt*1e3, x.numel()*4/dt/1e9))
t=time.perf_counter()
for _ in range(20): oh.copy_(o, non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/20
print("D2H 10MB: %.3f ms  %.1f GB/s" % (dt*1e3, o.numel()*4/dt/1e9))
PY
for mb in 16 24 48; do echo "slice_mb $mb"; TAC_HOST_SLICE_MB=$mb python bench.py --steps 200 --warmup 5 --cpu-seconds 0.1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['e2e']['ms_per_step'], d['e2e']['value'])"; done
