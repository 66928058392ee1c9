for mb in 200 1000 3000; do
  for c in 0 512 768 1024 1536 2048 3072 4096 6144 8192; do
    if [ $c = 0 ]; then unset CORNETTO_SDUST_CHUNK; else export CORNETTO_SDUST_CHUNK=$c; fi
    echo "mb=$mb chunk=$c $(python scripts/prof_sdust.py $mb 2 | tail -1)"
  done
done
