# Launch list of the driver's bench command shape (run under gpurun, one GPU): every kernel launch with its duration
# (cold-cache, serialised under ncu: shares, not absolutes).  The CPU legs and the batch / sweep legs are left out to keep
# the list to the headline grid and the one-rank leg of the 262 144-point grid.
TAG=${1:-r02s}
ARGS="--steps 20 --warmup 5 --no-cpu-baseline --batch-sets 0 --no-sweep"
mkdir -p gpurun_out
timeout 60 python bench.py $ARGS > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err; echo "plain exit $?"
timeout 80 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py $ARGS > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu exit $?"
python profiles/launch_summary.py gpurun_out/${TAG}_launches.csv | tee gpurun_out/${TAG}_launches.txt | cut -c1-170
