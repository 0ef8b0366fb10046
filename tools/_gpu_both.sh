mkdir -p gpurun_out/$TAG
for V in "" pair7 pair8; do
echo "variant=$V" | tee -a gpurun_out/$TAG/pairing_bench.txt
C12381_LIB_VARIANT=$V timeout 600 python tools/_pairing_bench.py 4 65536 2>&1 | tee -a gpurun_out/$TAG/pairing_bench.txt | tail -3
done
TAG=$TAG SKIP_NCU=1 bash tools/_gpu_quick.sh
