mkdir -p gpurun_out/$TAG
for V in "" $VARIANTS; do
  C12381_LIB_VARIANT=$V timeout 300 python tools/_pairing_bench.py 65536 4 2>&1 | tail -2 | tee -a gpurun_out/$TAG/pairing_bench.txt
done
C12381_PAIRING=scalar timeout 300 python tools/_pairing_bench.py 65536 4 2>&1 | tail -2 | tee -a gpurun_out/$TAG/pairing_bench.txt
