mkdir -p gpurun_out
H=$PWD/bodyfitting_b200/libbodyfit_b200_head.so
for i in 1 2; do
echo head; BODYFIT_LIB=$H timeout 300 python tools/time_kernels.py 2>&1 | tail -n 1
echo new; timeout 300 python tools/time_kernels.py 2>&1 | tail -n 1
done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ops.py -x -q -m gpu 2>&1 | tail -n 2
