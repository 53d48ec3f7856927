python -m pytest tests -m gpu -x -q > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_pytest.log; tail -4 gpurun_out/r02f_pytest.log
R=r02f bash tools/final_evidence.sh
