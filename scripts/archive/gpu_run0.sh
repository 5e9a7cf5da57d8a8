set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 300 bash tests/golden/make_golden.sh /tmp/ref_raw gpurun_out/golden > gpurun_out/golden.log 2>&1; echo "golden rc=$?"
tail -5 gpurun_out/golden.log
mkdir -p /tmp/drop_raw
for s in 7 8 5; do timeout 120 oracle/_ref/ref_host_on_psolver --scene $s --mode staged --steps 1 --out /tmp/drop_raw/scene$s >> gpurun_out/dropin.log 2>&1; echo "dropin $s rc=$?"; done
python tests/golden/pack_golden.py /tmp/drop_raw gpurun_out/golden_dropin >> gpurun_out/dropin.log 2>&1
tail -5 gpurun_out/dropin.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/pytest.log
timeout 300 oracle/_ref/ref_gpu --scene c3 --side 100 --grid 256 --max 1100000 --mode whole --steps 4 --dump-every 0 --out /tmp/ref_c3 2>&1 | tail -3
