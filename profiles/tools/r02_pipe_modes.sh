# SPT_PIPE_MODE=2 (two half-size waves per frame on a lane pair chosen by frame parity) existed only for this comparison and was removed afterwards
set -x
python -m pytest tests/test_gpu_parity.py -q -m gpu -k "two_frames or tile_sets" 2>&1 | tail -3
for m in 0 1 2; do echo "== SPT_PIPE_MODE=$m"; SPT_PIPE_MODE=$m python profiles/tools/pipelined_ranks.py; done 2>&1 | tee gpurun_out/r02_pipe_modes.log
