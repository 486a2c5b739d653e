cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export FNSSL_TC_WAIT_TIMEOUT=1 FNSSL_TC_PAIR_MIN=1
F="full_in16_H128x2 full_in256_H128x2_add narrow_in256_H128x2 narrow_in256+16_H128x2_add"
for v in t5flat t5split; do
  export FNSSL_B200_LIB=$GRAFT_REPO_ROOT/fn_ssl_b200/variants/libfnssl_b200_$v.so
  echo "=== $v"; timeout 300 python tools/tc5_debug.py 2>&1 | grep -c "rel-to-max err = [0-9.]*e-04" ; timeout 300 python tools/lstm_time.py $F | grep -v "_B1\|_B2"
done
