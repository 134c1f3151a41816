set -u
export MMC_SPARSE_DEVICE_MIN=0
T="tests/test_gpu_golden.py::test_cuda_sparse_rows_sorted_on_device tests/test_gpu_edge_cases.py::test_edge_cases_cuda_two_bit_seq"
echo "---- memcheck (goldens incl. test17a --insertions with the device-side sparse finalize forced, all edge cases with 2-bit SEQ transport)" > gpurun_out/sanitizer_r01e.txt
timeout 1200 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest $T -x -q 2>&1 | grep -E "=========|passed|failed" | tail -6 >> gpurun_out/sanitizer_r01e.txt
echo "---- racecheck (same tests)" >> gpurun_out/sanitizer_r01e.txt
timeout 1500 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest $T -x -q 2>&1 | grep -E "=========|passed|failed" | tail -6 >> gpurun_out/sanitizer_r01e.txt
cat gpurun_out/sanitizer_r01e.txt
