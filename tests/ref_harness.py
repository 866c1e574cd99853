"""TEST INFRASTRUCTURE ONLY: makes the REAL, unmodified ``torch_em`` package importable in this image.

``import torch_em`` needs imageio / skimage / bioimage_cpp / elf / kornia / h5py / ... which are absent here (no wheels,
no network).  None of them is *called* on the path under test (``default_segmentation_trainer`` with ``logger=None``,
in-memory tensors, no label transform), so a ``sys.meta_path`` finder serves empty stub modules for exactly the roots
that cannot be imported.  The package itself comes from ``/root/reference`` (build container) or from
``baseline/_ref`` (the offline ``pip install --no-deps --target baseline/_ref`` of the reference; git-ignored, travels to
the GPU box).  Nothing under ``torch-em_b200/`` imports this file.
"""
import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUB_ROOTS = ("imageio", "skimage", "bioimage_cpp", "elf", "kornia", "h5py", "matplotlib", "natsort", "tifffile", "mrcfile",
              "bioimageio", "zarr", "z5py", "nifty", "vigra", "affogato", "tensorboard", "xarray", "pooch", "cv2",
              "nibabel", "imagecodecs", "torchvision")
# NOT stubbed although absent: napari -- the reference guards it itself (``try: from napari.utils import progress as tqdm
# except ImportError: from tqdm import tqdm``, util/prediction.py:13-16) and a stub would shadow the real tqdm.


def reference_root():
    """Directory that contains the reference's ``torch_em`` package, or None."""
    for cand in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if os.path.isfile(os.path.join(cand, "torch_em", "__init__.py")):
            return cand
    return None


class _StubMeta(type):
    """Class attributes of a stub class (enum members such as ``kornia.constants.Resample.BILINEAR``) are plain strings."""

    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return f"{cls.__name__}.{name}"


class _StubModule(types.ModuleType):
    """Any attribute is another stub (module-like and callable-class-like), so ``from x.y import z`` always succeeds."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        full = f"{self.__name__}.{name}"
        if name[:1].isupper():
            val = _StubMeta(name, (), {"__module__": self.__name__, "__init__": lambda self, *a, **k: None})
        else:
            val = _StubModule(full)
            val.__path__ = []
            sys.modules.setdefault(full, val)
        setattr(self, name, val)
        return val

    def __call__(self, *a, **k):
        raise RuntimeError(f"stub of the absent third-party module {self.__name__} was called")


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def __init__(self, roots):
        self.roots = set(roots)

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in self.roots:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_installed = False


def import_torch_em():
    """Returns the real ``torch_em`` module (stubs served only for third-party roots that do not import), or None."""
    global _installed
    root = reference_root()
    if root is None:
        return None
    if not _installed:
        missing = []
        for r in STUB_ROOTS:
            if r in sys.modules:
                continue
            try:
                if importlib.util.find_spec(r) is None:
                    missing.append(r)
            except (ImportError, ValueError):
                missing.append(r)
        sys.meta_path.append(_StubFinder(missing))
        if root not in sys.path:
            sys.path.insert(0, root)
        _installed = True
    import torch_em
    return torch_em


def load_reference_module(name, rel):
    """One torch-only reference file (e.g. ``model/unet.py``) loaded by path, without importing the package."""
    root = reference_root()
    if root is None:
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, os.path.join(root, "torch_em", rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod
