mkdir -p gpurun_out/$TAG
timeout 300 python tools/_pairing_bench.py 17760 4 2>&1 | tail -2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pairing_coop -s 1 -c 1 -o gpurun_out/$TAG/prof_pairing_coop -f python tools/_pairing_bench.py 17760 4 > gpurun_out/$TAG/ncu_pair.log 2>&1
echo "ncu exit $?"
