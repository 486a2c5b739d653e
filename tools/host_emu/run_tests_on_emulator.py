"""Run `-m gpu` tests of tests/test_training_backward.py on the CPU against the HOST EMULATION of the training kernels
(tools/host_emu/common.cuh) -- a development aid for kernel / binding edits made without GPU time, NOT a product path and not part
of any test suite: it monkeypatches this process only (Tensor.cuda -> identity, the C-ABI loader -> the emulated library, the
CUDA-only guards off) and then calls the selected test functions with their real bodies, so the Python side (autograd Functions,
packing, ctypes signatures, module wiring) runs exactly as on the GPU while the kernels run as OS threads.

    python tools/host_emu/run_tests_on_emulator.py [--module test_training_backward | test_gpu_ipdnet2] [-k substring] ... [--oracle-frontend]

Tests that need kernels outside the emulated set (front end, tcgen05 engine) fail with "not emulated"; --oracle-frontend stands the
oracle's STFT / feature assembly in for the (TMA-staged, not emulatable) front-end kernels so that the training-module tests can
run through everything behind them."""
import contextlib
import ctypes as C
import inspect
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import build as emu_build  # noqa: E402
from fn_ssl_b200 import _lib, ops  # noqa: E402


class EmuLib:
    def __init__(self):
        self._so = C.CDLL(emu_build.build())

    def __getattr__(self, name):
        try:
            fn = getattr(self._so, name)
        except AttributeError:
            raise RuntimeError(f"{name}: not emulated (kernel outside tools/host_emu's source set)")
        res, args = _lib.SIGNATURES[name]
        fn.restype, fn.argtypes = res, args
        return fn


def install():
    emu = EmuLib()
    _lib.load = lambda build_if_missing=True: emu
    _lib.check = lambda rc: (_ for _ in ()).throw(RuntimeError("emulated call failed: " + emu._so.emu_last_error().decode())) if rc else None
    emu._so.emu_last_error.restype = C.c_char_p
    ops._need_cuda = lambda *t: None
    ops._stream = lambda: None
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    _tensor_to, _module_to = torch.Tensor.to, torch.nn.Module.to

    def _strip(args, kwargs):           # .to("cuda") / .to(device=...) -> stay where we are
        args = tuple(a for a in args if not (isinstance(a, (str, torch.device)) and str(a).startswith("cuda")))
        kwargs = {k: v for k, v in kwargs.items() if not (k == "device" and str(v).startswith("cuda"))}
        return args, kwargs

    def tensor_to(self, *a, **k):
        a, k = _strip(a, k)
        return _tensor_to(self, *a, **k) if (a or k) else self

    def module_to(self, *a, **k):
        a, k = _strip(a, k)
        return _module_to(self, *a, **k) if (a or k) else self

    torch.Tensor.to, torch.nn.Module.to = tensor_to, module_to
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.current_device = lambda: None            # == torch.device('cpu').index: on_tensor_device takes the direct path
    torch.cuda.device = lambda *a, **k: contextlib.nullcontext()
    # the train-mode modules insist on CUDA inputs: tell them every tensor is one (this process only)
    torch.Tensor.is_cuda = property(lambda self: True)


def install_oracle_frontend():
    from oracle import fnssl_oracle as orc

    def stft(signal, win_len=512, hop=256, nfft=512, want_magsum=False):
        return orc.stft(signal.float()), None

    def features(spec, magsum, pairing, norm, sample_length, eps, dtype, want_cfirst=False, mu=None):
        st = spec.permute(0, 3, 1, 2)
        reb = st if pairing == "ALL" else orc.add_ch_to_batch(st, pairing)
        if norm == ops.NORM_FORGETTING:
            m = orc.forgetting_norm(torch.abs(reb), sample_length)
            re, im = torch.real(reb) / (m + eps), torch.imag(reb) / (m + eps)
        else:
            m, re, im = None, torch.real(reb), torch.imag(reb)
        return None, m, torch.cat((re, im), dim=1)[:, :, 1:257, :].contiguous()

    ops.stft, ops.features = stft, features


class MonkeyPatch:
    def __init__(self):
        self.saved = {}

    def setenv(self, k, v):
        self.saved.setdefault(k, os.environ.get(k))
        os.environ[k] = v

    def undo(self):
        for k, v in self.saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def main():
    keys = [sys.argv[i + 1] for i, a in enumerate(sys.argv[:-1]) if a == "-k"]
    install()
    if "--oracle-frontend" in sys.argv:
        install_oracle_frontend()
    modname = sys.argv[sys.argv.index("--module") + 1] if "--module" in sys.argv else "test_training_backward"
    T = __import__(modname)
    g = np.load(os.path.join(ROOT, "tests", "golden", "grad_golden.npz"))
    gi = np.load(os.path.join(ROOT, "tests", "golden", "grad_ipdnet_golden.npz"))
    fixtures = {"golden_ipdnet2": "ipdnet2_golden.npz", "golden_fnssl": "fnssl_golden.npz", "golden_ipdnet": "ipdnet_golden.npz"}
    module_gpu = any(getattr(m, "name", "") == "gpu" for m in ([getattr(T, "pytestmark", None)] if not isinstance(getattr(T, "pytestmark", None), list) else T.pytestmark) if m is not None)
    failed = 0
    for name, fn in sorted(vars(T).items()):
        if not name.startswith("test_") or not callable(fn) or not (module_gpu or any(m.name == "gpu" for m in getattr(fn, "pytestmark", []))):
            continue
        if keys and not any(k in name for k in keys):
            continue
        cases = [{}]
        for m in getattr(fn, "pytestmark", []):
            if m.name == "parametrize":
                names = [n.strip() for n in m.args[0].split(",")]
                cases = [dict(c, **dict(zip(names, v if len(names) > 1 else (v,)))) for c in cases for v in m.args[1]]
        for case in cases:
            kwargs = dict(case)
            mp = MonkeyPatch()
            for p in inspect.signature(fn).parameters:
                if p == "g":
                    kwargs[p] = g
                elif p == "gi":
                    kwargs[p] = gi
                elif p == "monkeypatch":
                    kwargs[p] = mp
                elif p in fixtures:
                    kwargs[p] = np.load(os.path.join(ROOT, "tests", "golden", fixtures[p]))
            t0 = time.time()
            try:
                fn(**kwargs)
                status = "ok"
            except Exception as exc:      # noqa: BLE001
                status = f"FAILED: {type(exc).__name__}: {str(exc)[:300]}"
                failed += 1
            mp.undo()
            print(f"{name}{case if case else ''}: {status}  ({time.time() - t0:.1f} s)", flush=True)
    sys.exit(1 if failed else 0)


if __name__ == "__main__":
    main()
