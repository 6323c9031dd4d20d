D=lenslesspicam_b200/_dev
LPC_LIB_PATH=$D/g0.so python scripts_dev/chk_variant.py 2>&1 | tail -1
for v in g1 g2 g3; do LPC_LIB_PATH=$D/$v.so python scripts_dev/chk_variant.py 2>&1 | tail -1; done
for v in g0 g1 g2 g3 g0 g1 g2 g3; do
echo "### $v"
LPC_LIB_PATH=$D/$v.so scripts_dev/ab_fista.sh "X=1" | tail -1
done
