"""Test doubles for libganb200 (host-logic tests on a machine without a GPU).  They live here, not in the product
package: the product only offers kernels.install_test_double(lib) to inject one."""


class NullLib:
    """Every entry point is a no-op that reports success; size queries return a token size and the "is this shape
    supported" queries (`*_rows`, `*_supported`) answer no.  NO arithmetic is performed: outputs stay uninitialised."""

    def __getattr__(self, name):
        if name.endswith("_workspace"):
            return lambda *a: 16
        return lambda *a: 0


class RecordingLib:
    """Records (name, args) of every kernel call and reports success; pure queries are answered, not recorded."""

    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        if name.endswith("_workspace"):
            return lambda *a: 16
        if name.endswith("_rows") or name.endswith("_supported"):
            return lambda *a: 0
        if name == "ganb_launch_count":
            return lambda *a: len(self.calls)

        def fn(*a):
            self.calls.append((name, a))
            return 0
        return fn

    def names(self):
        return [c[0] for c in self.calls]


def install(lib=None):
    """Installs `lib` (default: a NullLib) as the library of gan_lib_tensorflow_b200.kernels; returns it."""
    from gan_lib_tensorflow_b200 import kernels as K

    lib = lib if lib is not None else NullLib()
    K.install_test_double(lib)
    return lib


def uninstall():
    from gan_lib_tensorflow_b200 import kernels as K

    K.install_test_double(None)
