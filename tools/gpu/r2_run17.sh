cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export FNSSL_TC_WAIT_TIMEOUT=0 FNSSL_TC_PAIR_MIN=1
F="full_in16_H128x2 full_in256_H128x2_add narrow_in256_H128x2"
echo "== tc5 flat, drain relay, cta-scope H_FREE"; timeout 200 python tools/lstm_time.py $F | grep -v "_B1\|_B2"
echo "== + cluster-scope H_FREE (debug 32)"; FNSSL_TC_DEBUG=32 timeout 200 python tools/lstm_time.py $F | grep -v "_B"
echo "== debug 8"; FNSSL_TC_DEBUG=8 timeout 200 python tools/lstm_time.py $F | grep -v "_B"
FNSSL_TC_WAIT_TIMEOUT=1 timeout 300 python tools/tc5_debug.py 2>&1 | grep -c "rel-to-max err = [0-9.]*e-04"
python tools/tc5_trace.py 2>&1 | head -17
