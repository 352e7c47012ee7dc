for n in 1024 4096; do
  for cfg in "16384 0" "16384 1" "16384 2" "16384 3" "32768 1" "8192 4"; do set -- $cfg
    timeout 10 stdbuf -oL ./tools/xchg_bench --n $n --tile $1 --load $2 --iters 3000 || echo "TIMEOUT/FAIL n=$n cfg=$cfg"
  done
done
