import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from test_gpu_kernels import _engine
dev = torch.device("cuda:0")
rows = 65536; n = rows * 128
eng, _lib, _stream = _engine("dopri5", torch.float32, n, 0.0371, 0.5, 1.0)
lib = eng.lib
W = (torch.randn(128, 128) * 0.09).to(dev)
planes = torch.empty(int(lib.tdq_linear_weights_bytes(128)), dtype=torch.uint8, device=dev)
_lib.check(lib.tdq_linear_prepare(0, W.data_ptr(), 128, planes.data_ptr(), _stream()))
NS = 3
sets = [[(torch.rand(n, device=dev) * 2 - 1) if os.environ.get("MB_UNIFORM") else torch.randn(n, device=dev) for _ in range(8)] for _ in range(NS)]      # y0, k0..k6
outs = [[torch.empty(n, device=dev) for _ in range(3)] for _ in range(NS)]
ctrl, tabp, dc = eng.ctrl.data_ptr(), C.byref(eng.tab), eng.dt_code
def launch(row, s):
    st = sets[s % NS]; o = outs[s % NS]
    kp = _lib.ptr_array([k.data_ptr() for k in st[1:]])
    last = row == 5
    _lib.check(lib.tdq_linear_stage(ctrl, tabp, dc, row, o[0].data_ptr(), o[1].data_ptr() if last else None,
                                    o[2].data_ptr() if last else None, st[0].data_ptr(), kp, planes.data_ptr(), 128, n, _stream()))
def launch_combine(row, s):
    st = sets[s % NS]; o = outs[s % NS]
    kp = _lib.ptr_array([k.data_ptr() for k in st[1:]])
    if row == 5:
        _lib.check(lib.tdq_stage_combine_final(ctrl, tabp, dc, o[1].data_ptr(), o[2].data_ptr(), st[0].data_ptr(), kp, n, _stream()))
    else:
        _lib.check(lib.tdq_stage_combine(ctrl, tabp, dc, row, o[0].data_ptr(), st[0].data_ptr(), kp, n, _stream()))
for name, fn in (("linear_stage", launch), ("stage_combine", launch_combine)):
    for row in range(6):
        for i in range(3):
            fn(row, i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(21):
            fn(row, i)
        e1.record(); torch.cuda.synchronize()
        print(name, "row", row, "us", e0.elapsed_time(e1) / 21 * 1e3)
# the sequence of one attempt, as the engine issues it (rows 0..5 on ONE set: producer -> consumer through L2)
def attempt(s):
    st = sets[s % NS]
    k = [st[1].data_ptr()] + [None] * 6
    for row in range(6):
        last = row == 5
        out = st[2 + row]
        _lib.check(lib.tdq_linear_stage(ctrl, tabp, dc, row, out.data_ptr(), outs[s % NS][1].data_ptr() if last else None,
                                        outs[s % NS][2].data_ptr() if last else None, st[0].data_ptr(), _lib.ptr_array(k),
                                        planes.data_ptr(), 128, n, _stream()))
        k[row + 1] = out.data_ptr()
for i in range(3):
    attempt(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(12):
    attempt(i)
e1.record(); torch.cuda.synchronize()
print("six fused rows in sequence: us per attempt", e0.elapsed_time(e1) / 12 * 1e3)
