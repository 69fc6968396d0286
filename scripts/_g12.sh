set +e
mkdir -p gpurun_out/r2k
timeout 900 python -m pytest tests/test_gpu_linear.py -q > gpurun_out/r2k/pytest_linear.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k/pytest_linear.log
tail -12 gpurun_out/r2k/pytest_linear.log
cat > /tmp/one.py <<'PY'
import sys, os, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import problems as P
import torchdiffeq_b200 as tdq
dev = torch.device('cuda:0')
f = tdq.LinearField(P.skew_matrix(128, torch.float32).to(dev))
y0 = torch.randn(65536, 128, generator=torch.Generator().manual_seed(1)).to(dev)
t = torch.tensor([0., 1.], device=dev)
with torch.no_grad():
    y = tdq.odeint(f, y0, t, method='dopri5', rtol=1e-5, atol=1e-7, options={"graph": False})
torch.cuda.synchronize()
print(tdq.last_stats())
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2k/launches_fused.csv python /tmp/one.py > gpurun_out/r2k/ncu_one.log 2>&1
tail -3 gpurun_out/r2k/ncu_one.log
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r2k/launches_fused.csv')) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ki][:90]; v = float(r[vi].replace(',', ''))
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print('%6d %10.1f us avg %8.2f  %s' % (c, v / 1e3, v / c / 1e3, k))
PY
