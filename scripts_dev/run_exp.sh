D=lenslesspicam_b200/_dev
export LPC_LIB_PATH=$D/p12.so
python scripts_dev/chk_variant.py 2>&1 | tail -1
LPC_B12_PLANES=1 python scripts_dev/chk_variant.py 2>&1 | tail -1
scripts_dev/ab.sh "LPC_B12_PLANES=0" "LPC_B12_PLANES=1" "LPC_B12_PLANES=0" "LPC_B12_PLANES=1" "LPC_B12_PLANES=2"
