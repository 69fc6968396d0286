set +e
mkdir -p gpurun_out/r2k
timeout 300 python scripts/_mb_linear.py 2>&1 | grep "linear_stage\|six" > gpurun_out/r2k/mb2.log; cat gpurun_out/r2k/mb2.log
timeout 600 python -m pytest tests/test_gpu_linear.py -q -x 2>&1 | tail -3
timeout 300 python - > gpurun_out/r2k/quick2.log 2>&1 <<'PY'
import sys, os, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import problems as P
import torchdiffeq_b200 as tdq
dev = torch.device('cuda:0')
f = tdq.LinearField(P.skew_matrix(128, torch.float32).to(dev))
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
    print(opts, 'ms/solve', e0.elapsed_time(e1) / 5, tdq.last_stats())
PY
cat gpurun_out/r2k/quick2.log | tail -3
