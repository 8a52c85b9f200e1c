# Round-2 batch M: launch lists (time + DRAM bytes) of the analytic-sphere scene and the fog + smoke scene.
set -x
O=gpurun_out
for sc in spheres media; do
  python tools/prof_path.py --scene $sc --spp 8 > $O/r02m_prof_${sc}_plain.log 2>&1 || exit 1
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r02m_${sc}_launches.csv python tools/prof_path.py --scene $sc --spp 8 > $O/r02m_${sc}_ncu.log 2>&1
  python tools/launch_share.py $O/r02m_${sc}_launches.csv k_raygen | tee $O/r02m_${sc}_share.txt
done
