mkdir -p gpurun_out/$TAG
timeout 600 python tools/_pairing_bench.py 4 24576,32768,37888,49152,65536 2>&1 | tee gpurun_out/$TAG/pairing_bench.txt | tail -16
TAG=$TAG SKIP_NCU=1 bash tools/_gpu_quick.sh
