"""Names only (VMamba.py:14 imports them at module scope; they are used by .flops() which goldens never call)."""


def _unavailable(*_a, **_k):
    raise RuntimeError("fvcore is not installed; shim provides names only")


FlopCountAnalysis = flop_count_str = flop_count = parameter_count = _unavailable
