P=physics-informed-neural-network-for-explainable-fault-diagnosis-in-fuel-cells_b200
for v in 1 E F G H 1; do
  echo "== variant $v"
  B200PINN_LIB=$P/build/res$v/libb200pinn.so timeout 120 python profiles/wide_res_check.py quick 2>&1 | tail -1
  B200PINN_LIB=$P/build/res$v/libb200pinn.so timeout 120 python profiles/wide_res_check.py time 2>&1 | tail -2
done
