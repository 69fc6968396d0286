set +e
mkdir -p gpurun_out/r2k
timeout 900 python -m pytest tests/test_gpu_linear.py -x -q > gpurun_out/r2k/pytest_linear.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k/pytest_linear.log
tail -15 gpurun_out/r2k/pytest_linear.log
timeout 300 python - > gpurun_out/r2k/quick.log 2>&1 <<'PY'
import sys, torch, time
sys.path.insert(0, 'tests')
import problems as P
import torchdiffeq_b200 as tdq
dev = torch.device('cuda:0')
A = P.skew_matrix(128, torch.float32).to(dev)
f = tdq.LinearField(A)
y0 = torch.randn(65536, 128, generator=torch.Generator().manual_seed(1)).to(dev)
t = torch.tensor([0., 10.], device=dev)
for opts in ({}, {"fused_linear": False}):
    with torch.no_grad():
        for _ in range(3):
            y = tdq.odeint(f, y0, t, method='dopri5', rtol=1e-5, atol=1e-7, options=dict(opts))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            y = tdq.odeint(f, y0, t, method='dopri5', rtol=1e-5, atol=1e-7, options=dict(opts))
        e1.record(); torch.cuda.synchronize()
    print(opts, 'ms/solve', e0.elapsed_time(e1) / 5, tdq.last_stats(), float(y[-1].norm(dim=1).sub(y0.norm(dim=1)).abs().max()))
PY
cat gpurun_out/r2k/quick.log | tail -5
