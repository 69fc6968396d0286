"""Stand-in for the CALLER side of the reference's solver-registry seam, used by the plug-in tests when the
reference package itself is not importable (it normally is: baseline/_ref travels with the repo).

It restates only what reaches a solver through the seam (misc.py:200-345 _check_inputs and odeint.py:90-108):
tuple states flattened with torch.cat, time made ascending with a sign-flipping wrapper, the perturb= wrapper, the
callback attributes (with a null lambda for absent ones), options['norm'] always set (the module-level _rms_norm for
tensor states, an anonymous closure for tuple states) and the SOLVERS[method](func=..., y0=..., rtol=..., atol=...,
**options).integrate(t) / .integrate_until_event(t0, event_fn) calls.  Test infrastructure, not product."""
import torch

_all_callback_names = ['callback_step', 'callback_accept_step', 'callback_reject_step']
_null_callback = lambda *args, **kwargs: None
SOLVERS = {}


def _rms_norm(tensor):
    return tensor.abs().pow(2).mean().sqrt()


def _mixed_norm(tensor_tuple):
    return max([_rms_norm(tensor) for tensor in tensor_tuple])


def _flat_to_shape(tensor, length, shapes):
    out, total = [], 0
    for shape in shapes:
        nxt = total + shape.numel()
        out.append(tensor[..., total:nxt].view((*length, *shape)))
        total = nxt
    return tuple(out)


class _TupleFunc(torch.nn.Module):
    def __init__(self, base_func, shapes):
        super().__init__()
        self.base_func, self.shapes = base_func, shapes

    def forward(self, t, y):
        f = self.base_func(t, _flat_to_shape(y, (), self.shapes))
        return torch.cat([f_.reshape(-1) for f_ in f])


class _ReverseFunc(torch.nn.Module):
    def __init__(self, base_func, mul=1.0):
        super().__init__()
        self.base_func, self.mul = base_func, mul

    def forward(self, t, y):
        return self.mul * self.base_func(-t, y)


class _PerturbFunc(torch.nn.Module):
    def __init__(self, base_func):
        super().__init__()
        self.base_func = base_func

    def forward(self, t, y, *, perturb=None):
        return self.base_func(t.to(y.abs().dtype), y)


def odeint(func, y0, t, *, rtol=1e-7, atol=1e-9, method=None, options=None, event_fn=None):
    original_func = func
    shapes = None
    if not isinstance(y0, torch.Tensor):
        shapes = [y_.shape for y_ in y0]
        y0 = torch.cat([y_.reshape(-1) for y_ in y0])
        func = _TupleFunc(func, shapes)
        if event_fn is not None:
            ev_user = event_fn
            event_fn = lambda t_, y_: ev_user(t_, _flat_to_shape(y_, (), shapes))
    options = {} if options is None else options.copy()
    method = method or 'dopri5'
    if shapes is not None:
        norm = options.get('norm', _mixed_norm)

        def _norm(tensor):
            return norm(_flat_to_shape(tensor, (), shapes))
        options['norm'] = _norm
    elif 'norm' not in options:
        options['norm'] = _rms_norm
    t_is_reversed = len(t) > 1 and bool(t[0] > t[1])
    if t_is_reversed:
        t = -t
        func = _ReverseFunc(func, mul=-1.0)
        if event_fn is not None:
            event_fn = _ReverseFunc(event_fn)
        for name in ('step_t', 'jump_t'):
            if name in options:
                options[name] = -options[name]
    func = _PerturbFunc(func)
    for name in _all_callback_names:
        cb = getattr(original_func, name, None)
        if cb is None:
            setattr(func, name, _null_callback)
        else:
            if shapes is not None:
                cb = (lambda t0, y_, dt, _cb=cb: _cb(t0, _flat_to_shape(y_, (), shapes), dt))
            if t_is_reversed:
                cb = (lambda t0, y_, dt, _cb=cb: _cb(-t0, y_, dt))
            setattr(func, name, cb)
    SOLVERS[method].valid_callbacks()
    solver = SOLVERS[method](func=func, y0=y0, rtol=rtol, atol=atol, **options)
    if event_fn is None:
        solution = solver.integrate(t)
    else:
        event_t, solution = solver.integrate_until_event(t[0], event_fn)
        event_t = event_t.to(t)
        if t_is_reversed:
            event_t = -event_t
    if shapes is not None:
        solution = _flat_to_shape(solution, (len(t),), shapes)
    return solution if event_fn is None else (event_t, solution)
