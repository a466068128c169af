"""Stand-ins for the C-ABI library and the coordinate manager so that the facade's HOST logic (lazy convs, fused
autograd nodes, weight-operand caching, stream fork/join bookkeeping) can run on CPU tensors without a GPU: every entry
point returns LGS_OK and does no arithmetic, kernel maps are zero tables of the right shapes."""
import torch

from languagegroundedsemseg_b200 import minkowski as E


class StubLib:
    def __init__(self):
        self.calls = {}

    def __getattr__(self, name):
        def f(*a):
            self.calls[name] = self.calls.get(name, 0) + 1
            return 1 if name.endswith("_supported") else 0
        return f


class FakeEvent:
    def record(self, s=None):
        pass

    def wait(self, s=None):
        pass


class FakeStream:
    cuda_stream = 0

    def wait_event(self, e):
        pass


class FakeManager:
    D = 3

    def __init__(self, sizes):
        self.sizes = sizes          # rows per tensor stride
        self.cache = {}

    def size(self, key):
        return self.sizes[key.tensor_stride[0]]

    def conv_maps(self, in_key, ks, stride, dil, transpose):
        ts = in_key.tensor_stride[0]
        out_ts = ts // stride if transpose else ts * stride
        ck = (ts, out_ts, ks)
        if ck not in self.cache:
            km = E.KernelMap()
            km.K, km.n_in, km.n_out = ks ** 3, self.sizes[ts], self.sizes[out_ts]
            km.fwd_table = torch.zeros((km.K, km.n_out), dtype=torch.int32)
            km.bwd_table = torch.zeros((km.K, km.n_in), dtype=torch.int32)
            km.bwd_reverse, km.counts = ts == out_ts, torch.zeros(km.K, dtype=torch.int32)
            self.cache[ck] = (E.CoordinateMapKey([out_ts] * 3), km)
        return self.cache[ck]


def install(setattr_fn, real_library=False):
    """patch the facade through `setattr_fn(obj, name, value)` (pytest's monkeypatch.setattr or plain setattr).
    real_library=True keeps the real C library behind the facade (to be used inside `_lib.trace()`, where the compute
    entry points record their arguments and return without touching a GPU); only the CUDA stream / scratch plumbing of
    the facade is replaced."""
    from languagegroundedsemseg_b200 import _lib
    stub = None
    if not real_library:
        stub = StubLib()
        setattr_fn(_lib, "load", lambda: stub)
    setattr_fn(E, "_stream", lambda: None)
    scratch = E._Scratch(torch.device("cpu"))
    setattr_fn(E, "_scratch64", lambda idx: scratch)
    setattr_fn(E, "_side_stream", lambda idx: (FakeStream(), FakeEvent(), FakeEvent()))
    setattr_fn(E, "_cur_stream_obj", lambda idx: FakeStream())
    return stub


def sparse_input(rows, channels, mgr):
    return E.SparseTensor._make(torch.randn(rows, channels), E.CoordinateMapKey([1, 1, 1]), mgr)
