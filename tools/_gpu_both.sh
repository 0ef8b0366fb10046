mkdir -p gpurun_out/$TAG
timeout 600 python tools/_msm_bench.py 2>&1 | tee gpurun_out/$TAG/msm_ab.txt | tail -16
TAG=$TAG SKIP_NCU=1 bash tools/_gpu_quick.sh
