# last-resort launch for lists beyond CAP_BIG: the new test, plain and under memcheck; the tests around it
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "last_resort or many_layers or seeded_2d or errors_are" 2>&1 | tail -15 | tee gpurun_out/r2cf_pytest.log
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "last_resort" > gpurun_out/r2cf_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/r2cf_memcheck.log
