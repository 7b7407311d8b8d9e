for nw in 4 8; do
  echo "v3 nw $nw: $(KE_PHASH_NW=$nw python tools/profile_kernels.py --reps 3 --only phash --images 8192 2>&1 | tail -1)"
done
KE_PHASH_NW=4 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "phash" 2>&1 | tail -3
