"""Build an A/B variant of the library: python profiles/build_variant.py <name> [-DFLAG ...] -> build/<name>/libb200pinn.so
(use with B200PINN_LIB=<that path> python profiles/quick_time.py ...)."""
import importlib.util, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "physics-informed-neural-network-for-explainable-fault-diagnosis-in-fuel-cells_b200")
spec = importlib.util.spec_from_file_location("b", os.path.join(PKG, "build.py"))
m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
D = os.path.join(PKG, "build", sys.argv[1])
os.makedirs(D, exist_ok=True)
print(m.build(force=True, extra_flags=sys.argv[2:], out=os.path.join(D, "libb200pinn.so"), objdir=D))
