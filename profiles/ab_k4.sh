# A/B of 64-wide MC kernel build variants (profiles/build_variant.py <name> -D...): timing
P=physics-informed-neural-network-for-explainable-fault-diagnosis-in-fuel-cells_b200
for v in "$@"; do
  echo "== variant $v"
  if [ $v = main ]; then unset B200PINN_LIB; else export B200PINN_LIB=$P/build/$v/libb200pinn.so; fi
  timeout 300 python profiles/quick_time.py 1000000 mc,fwd 2>&1 | tail -6
done
