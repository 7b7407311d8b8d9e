for n in 300000 700000 1000000; do
  echo "n=$n"; KE_JOIN_DEBUG=1 python tools/join_modes.py $n 2>&1 | grep -E "hybrid|ke_join|popc:|sliced:" | head -12
done
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "join" 2>&1 | tail -3
