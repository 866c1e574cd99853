"""Importable name of the ``torch-em_b200/`` package (a hyphen is not a legal module name).

``import torch_em_b200`` resolves sub-modules from ``<repo>/torch-em_b200/`` -- e.g.
``torch_em_b200.model.unet.UNet3d`` is ``torch-em_b200/model/unet.py`` -- so checkpoints written by
torch-em's ``DefaultTrainer`` (which pickles ``f"{cls.__module__}.{cls.__name__}"``, default_trainer.py:484-501)
can re-import the model class by dotted path.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "torch-em_b200")
__path__.insert(0, _real)

from ._api import *  # noqa: E402,F401,F403
from ._api import __all__  # noqa: E402,F401
