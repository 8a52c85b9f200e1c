# Round-2 final batch (2): smoke(), whole GPU suite, the default bench line and the reference arm of the current tree.
set -x
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02at_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r02at_smoke.log
python -m pytest tests -m gpu -x -q > $O/r02at_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02at_pytest.log
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02at_bench_ref.json 2> $O/r02at_bench_ref.err; echo "ref rc=$?" >> $O/r02at_bench_ref.err
python bench.py > $O/r02at_bench.json 2> $O/r02at_bench.err; echo "bench rc=$?" >> $O/r02at_bench.err
tail -2 $O/r02at_smoke.log; tail -3 $O/r02at_pytest.log; tail -2 $O/r02at_bench_ref.err; tail -2 $O/r02at_bench.err
