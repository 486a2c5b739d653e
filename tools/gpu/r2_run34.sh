cd $GRAFT_REPO_ROOT; O=gpurun_out; mkdir -p $O
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:lstm_tc --csv --log-file $O/r2_launches_34_layers.csv python tools/lstm_time.py full_in16_H128x2 narrow_in256_H128x2_add narrow_in256_H256x1_add > $O/r2_ncu_34.log 2>&1
grep -v "^==" $O/r2_launches_34_layers.csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
h=rows[0]; k=h.index('Kernel Name'); v=h.index('Metric Value'); g=h.index('Grid Size')
for r in rows[1:]: print(r[k][:40], r[g], float(r[v])/1e6, 'ms')
" | awk '{print}' | sort | uniq -c | sort -k1,1nr | head -40
grep "ms " $O/r2_ncu_34.log | head -20
