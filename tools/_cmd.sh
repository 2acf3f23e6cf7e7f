cd /root/repo; export PYTHONUNBUFFERED=1
for v in hg16 hg32 hg48 default; do
  if [ $v = default ]; then unset AKI_MMA_LIB; else export AKI_MMA_LIB=$PWD/build/libaki_$v.so; fi
  timeout 100 python tools/fwd_time.py 2>&1 | tail -1
  timeout 100 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:attn_fwd_sm100 -s 2 -c 1 python tools/fwd_only.py 2>&1 | grep -E "dram__bytes" | awk '{print "   ", $1, $2, $3}'
done
