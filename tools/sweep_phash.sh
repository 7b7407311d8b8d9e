for cfg in 1,8,1 2,8,1 1,8,2; do
 for dbg in 0 1 2 4 5; do
  echo "cfg $cfg dbg $dbg: $(KE_PHASH_DBG=$dbg KE_PHASH_CFG=$cfg python tools/profile_kernels.py --reps 2 --only phash --images 8192 2>&1 | tail -1)"
 done
done
