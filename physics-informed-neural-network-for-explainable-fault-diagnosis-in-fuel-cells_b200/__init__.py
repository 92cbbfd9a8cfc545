"""b200pinn -- B200-native PINN hot path (stack-voltage DNN forward/backward, fused
multi-physics residual loss, MC-dropout sweep) behind the reference's Python surface.

Public names mirror ``01_train_pinn_multiphysics_model.py``: ``DNN`` (01:389),
``PhysicsInformedNN`` (01:441), ``get_MC_samples`` (01:1413).  See DESIGN.md.
"""
from .nn import DNN, inject_masks  # noqa: F401
from .pinn import PhysicsInformedNN, LAMBDA_NAMES  # noqa: F401
from .mc import get_MC_samples, mc_dropout_device  # noqa: F401
from .export import create_comprehensive_results_array_v2, create_fault_labels, export_rows_device  # noqa: F401
from . import rf  # noqa: F401
from . import gmm  # noqa: F401
from .gmm import fit_gmm_and_get_probabilities  # noqa: F401
from .dropin import install  # noqa: F401

__all__ = ["DNN", "PhysicsInformedNN", "get_MC_samples", "mc_dropout_device", "inject_masks", "install",
           "LAMBDA_NAMES"]
