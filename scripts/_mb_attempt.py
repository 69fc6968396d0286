"""Microbenchmark of tdq_linear_attempt (csrc/tdq_attempt.cu) against the launches it replaces (6 x tdq_linear_stage +
tdq_error_norm_commit) at the configs[1] size, CUDA events, three rotating operand sets (201 MB each: larger than L2)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from test_gpu_kernels import _engine
dev = torch.device("cuda:0")
rows = int(os.environ.get("MB_ROWS", 65536)); n = rows * 128
method = os.environ.get("MB_METHOD", "dopri5")
eng, _lib, _stream = _engine(method, torch.float32, n, 0.0371, 0.5, 1.0)
S = eng.S
lib = eng.lib
W = (torch.randn(128, 128) * 0.09).to(dev)
planes = torch.empty(int(lib.tdq_linear_weights_bytes(128)), dtype=torch.uint8, device=dev)
_lib.check(lib.tdq_linear_prepare(0, W.data_ptr(), 128, planes.data_ptr(), _stream()))
NS = 3
sets = [[torch.randn(n, device=dev) for _ in range(2)] for _ in range(NS)]            # y0, k0
ks = [torch.empty(n, device=dev) for _ in range(S)]
y1, er = torch.empty(n, device=dev), torch.empty(n, device=dev)
ctrl, tabp, dc = eng.ctrl.data_ptr(), C.byref(eng.tab), eng.dt_code
kp_out = _lib.ptr_array([None] + [k.data_ptr() for k in ks])

def attempt_whole(s, store=0, fold=True):
    st = sets[s % NS]
    _lib.check(lib.tdq_linear_attempt(ctrl, tabp, dc, kp_out, y1.data_ptr(), er.data_ptr(), st[0].data_ptr(), st[1].data_ptr(),
                                      planes.data_ptr(), 128, n, eng.partials.data_ptr() if fold else None,
                                      eng.norm_out.data_ptr() if fold else None, None, store, _stream()))

def attempt_stages(s):
    st = sets[s % NS]
    k = [st[1].data_ptr()] + [None] * S
    for row in range(S):
        last = row == S - 1
        _lib.check(lib.tdq_linear_stage(ctrl, tabp, dc, row, ks[row].data_ptr(), y1.data_ptr() if last else None,
                                        er.data_ptr() if last else None, st[0].data_ptr(), _lib.ptr_array(k), planes.data_ptr(),
                                        128, n, _stream()))
        k[row + 1] = ks[row].data_ptr()
    _lib.check(lib.tdq_error_norm_commit(ctrl, dc, er.data_ptr(), ks[S - 1].data_ptr(), st[0].data_ptr(), y1.data_ptr(), None, None,
                                         None, 0, 0, 1, n, eng.partials.data_ptr(), eng.norm_out.data_ptr(), None, _stream()))

def timeit(name, fn, reps=15):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print("%-58s %9.1f us" % (name, us), flush=True)
    return us

# reference for a bitwise check of every variant
attempt_stages(0)
torch.cuda.synchronize()
ref = [k.clone() for k in ks] + [y1.clone(), er.clone(), eng.norm_out.clone()]
for accs in ("", "1"):
    if accs:
        os.environ["TDQ_ATTEMPT_ACCS2"] = "1"
    else:
        os.environ.pop("TDQ_ATTEMPT_ACCS2", None)
    for k in ks:
        k.zero_()
    attempt_whole(0, 1, True)
    torch.cuda.synchronize()
    got = [k.clone() for k in ks] + [y1.clone(), er.clone(), eng.norm_out.clone()]
    ok = all(torch.equal(a_, b_) for a_, b_ in zip(got[:-1], ref[:-1]))
    rel = max(float((a_ - b_).abs().max() / b_.abs().max()) for a_, b_ in zip(got[:-1], ref[:-1]))
    print("split accumulators" if accs else "one accumulator pair", "bitwise", ok, "max rel diff", rel, "norm rel diff",
          float((got[-1][0] - ref[-1][0]).abs() / ref[-1][0]))
    timeit("%s whole attempt [%s], norm folded, stages not stored" % (method, "2 x (big, small)" if accs else "big, small"),
           lambda i: attempt_whole(i, 0, True))
os.environ.pop("TDQ_ATTEMPT_ACCS2", None)
a = timeit("%s whole attempt, norm folded, stages not stored" % method, lambda i: attempt_whole(i, 0, True))
b = timeit("%s whole attempt, norm folded, stages stored" % method, lambda i: attempt_whole(i, 1, True))
c = timeit("%s whole attempt, no norm, stages stored" % method, lambda i: attempt_whole(i, 1, False))
d = timeit("%s %d x tdq_linear_stage + tdq_error_norm_commit" % (method, S), attempt_stages)
flops = 2.0 * rows * 128 * 128 * 6 * S
print("whole attempt: %.1f TFLOP/s of bf16 products, %.2f TB/s of the 4 N s algorithmic bytes; speed-up %.2fx"
      % (flops / a / 1e6, 4.0 * n * 4 / a / 1e6, d / a))
