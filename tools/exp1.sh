mkdir -p gpurun_out
echo "== builtin (nvcc AOT)"; python tools/sweep_gpu.py --fp 8 --sizes 135,441,126,128 --check 0 2>&1 | cut -c1-120
python tools/sweep_gpu.py --fp 4 --sizes 480,500 --check 0 2>&1 | cut -c1-120
echo "== JIT (NVRTC)"; BBFFT_CUDA_NO_BUILTIN=1 python tools/sweep_gpu.py --fp 8 --sizes 135,441,126,128 --check 0 2>&1 | cut -c1-120
BBFFT_CUDA_NO_BUILTIN=1 python tools/sweep_gpu.py --fp 4 --sizes 480,500 --check 0 2>&1 | cut -c1-120
