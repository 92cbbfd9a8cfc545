"""Rebind the reference module's names to this package ("how it drops in").

``01_train_pinn_multiphysics_model.py`` resolves ``DNN`` (01:464) and ``get_MC_samples``
(01:1914) from its module globals at call time, so after loading it by path::

    ref = load_reference_module(".../01_train_pinn_multiphysics_model.py")
    b200pinn.install(ref, level="C")

its own ``__main__`` body / helper functions drive the sm_100a kernels unchanged.

* level "A": ``ref.DNN`` -> ours (reference loops + autograd, kernels K1/K2);
* level "B": also ``ref.get_MC_samples`` -> ours (kernel K4);
* level "C": also ``ref.PhysicsInformedNN`` -> ours (device-resident scalers, K3, fused trainers)
  and ``ref.create_comprehensive_results_array_v2`` -> the device row writer (K5).
"""
from __future__ import annotations


def install(ref_module, level: str = "C"):
    from .nn import DNN
    from .mc import get_MC_samples
    from .pinn import PhysicsInformedNN

    level = level.upper()
    if level not in ("A", "B", "C"):
        raise ValueError("level must be 'A', 'B' or 'C'")
    ref_module.DNN = DNN
    if level in ("B", "C"):
        ref_module.get_MC_samples = get_MC_samples
    if level == "C":
        from .export import create_comprehensive_results_array_v2, create_fault_labels

        ref_module.PhysicsInformedNN = PhysicsInformedNN
        ref_module.create_comprehensive_results_array_v2 = create_comprehensive_results_array_v2
        ref_module.create_fault_labels = create_fault_labels
    return ref_module
