"""A/B aid: run the scaling experiment with the v1 library (tools/_v1/libv1.so) instead of the current one."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import b200nmpc
b200nmpc._ffi.LIB_PATH = Path(__file__).resolve().parent / '_v1' / 'libv1.so'
b200nmpc._ffi.EXPORTS = [e for e in b200nmpc._ffi.EXPORTS if e != 'nmpc_measure_fp64_peak']
import ctypes as C
_orig = C.CDLL
class _L(_orig):
    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            if name == 'nmpc_measure_fp64_peak':
                f = lambda *a: 1
                f.argtypes = None; f.restype = None
                class F:  # dummy
                    argtypes = None; restype = None
                    def __call__(self, *a): return 1
                return F()
            raise
C.CDLL = _L
sys.argv = ['gpu_debug.py', 'scaling']
import runpy
runpy.run_path(str(Path(__file__).resolve().parent / 'gpu_debug.py'), run_name='__main__')
