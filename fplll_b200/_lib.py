"""ctypes loader for the C-ABI libraries.  Fails loudly: there is no CPU or eager fallback."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
_cache = {}


class B200Error(RuntimeError):
    pass


def load(name):
    if name in _cache:
        return _cache[name]
    path = os.path.join(HERE, os.environ.get("B200_LIB_DIR", "lib"), name)
    if not os.path.exists(path):
        raise B200Error("%s is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(the product path has no CPU fallback)" % path)
    lib = C.CDLL(path)
    _cache[name] = lib
    return lib
