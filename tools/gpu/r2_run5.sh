cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out
export FNSSL_TC_WAIT_TIMEOUT=1
export FNSSL_B200_LIB=$GRAFT_REPO_ROOT/fn_ssl_b200/variants/libfnssl_b200_pub1.so
echo "== B1 shapes, default rows"; timeout 300 python tools/lstm_time.py B1 B2
echo "== B1 shapes, FNSSL_TC_ROWS=64"; FNSSL_TC_ROWS=64 timeout 300 python tools/lstm_time.py B1 B2
ncu --set full --clock-control none --import-source on -k regex:lstm_tc4 -s 4 -c 1 -o $O/r2_prof_pub1_in16 python tools/lstm_time.py full_in16_H128x2 > $O/r2_ncu_pub1_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lstm_tc4 -s 4 -c 1 -o $O/r2_prof_pub1_in256 python tools/lstm_time.py full_in256_H128x2_add > $O/r2_ncu_pub1_b.log 2>&1
ls -la $O/r2_prof_pub1*
