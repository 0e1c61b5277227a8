out=gpurun_out/r02zz; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $out/pytest.log; cat $out/pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"; cut -c1-200 $out/bench.json
