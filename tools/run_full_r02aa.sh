# Round-2 batch AA: whole GPU suite + the default bench line of the current tree.
set -x
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02aa_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02aa_pytest.log
python bench.py > $O/r02aa_bench.json 2> $O/r02aa_bench.err; echo "bench rc=$?" >> $O/r02aa_bench.err
tail -3 $O/r02aa_pytest.log; tail -3 $O/r02aa_bench.err
