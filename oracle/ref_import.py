"""Import shim for the *real* reference (baldhat/yolov10-3D, read-only at /root/reference).

TEST INFRASTRUCTURE ONLY.  Used by ``tests/golden/make_golden.py`` (run in the build
container, where /root/reference exists) to generate the committed golden fixtures and to
validate the CPU restatement in ``oracle/``.  Nothing on the product path, and nothing that
runs on the GPU box, imports this module: /root/reference does not exist there.

Recipe follows SURVEY.md §8(c): stub the four unused-but-imported third-party modules,
force numba's CUDA simulator (kitti_eval.py:248 eager-compiles a device function at import),
and point YOLO_CONFIG_DIR at a writable tmp dir.  No reference file is modified or copied.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("Y3D_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ultralytics"))


class _Anything:
    """Inert placeholder: any attribute / call / subscript yields another placeholder."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __getitem__(self, k):
        return _Anything()

    def __iter__(self):
        return iter(())


class _StubModule(types.ModuleType):
    """Module stub: unknown attributes resolve to inert placeholders / sub-stubs."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        full = f"{self.__name__}.{name}"
        if full in sys.modules:
            return sys.modules[full]
        return _Anything


class _StubFinder:
    """Resolves ``import stubbed_pkg.anything`` to a fresh stub module."""

    def __init__(self, roots):
        self.roots = tuple(roots)

    def find_spec(self, fullname, path=None, target=None):
        import importlib.machinery

        if fullname.split(".")[0] in self.roots:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def import_reference():
    """Returns the imported ``ultralytics`` package of the reference."""
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    os.environ.setdefault("NUMBA_ENABLE_CUDASIM", "1")
    os.environ.setdefault("YOLO_CONFIG_DIR", "/tmp/y3d_yolo_cfg")
    os.makedirs(os.environ["YOLO_CONFIG_DIR"], exist_ok=True)
    missing = []
    for name in ("matplotlib", "seaborn", "notion_client", "idlelib"):
        try:
            __import__(name)
        except Exception:
            missing.append(name)
    if missing:
        sys.meta_path.append(_StubFinder(missing))
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import ultralytics  # noqa: F401

    return ultralytics
