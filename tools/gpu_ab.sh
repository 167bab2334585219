set -u
for v in "" "AVSR_B200_LIB=/root/repo/avsr_tf1_b200/lib/libavsr_b200_old.so" "" "AVSR_B200_LIB=/root/repo/avsr_tf1_b200/lib/libavsr_b200_old.so"; do
  echo "--- $v"; env $v timeout -s KILL 200 python tools/ap_time.py 2>&1 | tail -2
done
