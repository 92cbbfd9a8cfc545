# A/B of build variants (profiles/build_variant.py <name> -D...): parity check + timing of the resident wide kernel
P=physics-informed-neural-network-for-explainable-fault-diagnosis-in-fuel-cells_b200
for v in "$@"; do
  echo "== variant $v"
  if [ $v = main ]; then unset B200PINN_LIB; else export B200PINN_LIB=$P/build/$v/libb200pinn.so; fi
  timeout 120 python profiles/wide_res_check.py quick 2>&1 | tail -1
  timeout 120 python profiles/wide_res_check.py time 2>&1 | tail -2
done
