mkdir -p gpurun_out/$TAG
for V in pc2 pc2i; do
echo "variant=$V" | tee -a gpurun_out/$TAG/ab.txt
C12381_LIB_VARIANT=$V timeout 600 python tools/_pairing_bench.py 4 1,16384,65536 2>&1 | tee -a gpurun_out/$TAG/ab.txt | tail -6
done
