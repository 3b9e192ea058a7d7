#!/bin/bash
# apply the experimental epilogue patch, run the GPU check, restore the tested tree if the GPU was not available
cd /root/repo
git apply scratch/${PATCH:-epilogue_combined.patch} || exit 1
(cd nvmolkit_b200/csrc && make -j8 2>&1 | grep -E "error" | head -3)
/usr/local/graft/bin/gpurun --timeout 500 -- 'timeout 150 python -m pytest tests/test_path_a_gpu.py tests/test_golden_fixtures.py -m gpu -q -x 2>&1 | tail -3; B200_TENSOR_CLUSTER=2 timeout 60 python -m pytest tests/test_path_a_gpu.py -m gpu -q -x -k "tensor" 2>&1 | tail -2; B200_TENSOR_CLUSTER=0 timeout 60 python -m pytest tests/test_path_a_gpu.py -m gpu -q -x -k "tensor" 2>&1 | tail -2; timeout 100 python bench.py --etkdg-mols 0 --cross-n 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d[\"value\"], d[\"ms_per_step\"], d[\"phases_ms\"][\"neighbor_pass_tc\"], d[\"parity_on_sample\"], d[\"n_clusters\"])"' > /tmp/try.log 2>&1
tail -10 /tmp/try.log
if grep -q "status=transient" /tmp/try.log; then
  git checkout nvmolkit_b200/csrc/tanimoto_tc.cu
  (cd nvmolkit_b200/csrc && make -j8 2>&1 | grep -E "error" | head -3)
  echo "RESTORED tested tree"
fi
