timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "klt" 2>&1 | tail -2
for i in 1 2; do
for lib in ab_old/libftk_old.so feature_tracker_b200/libftk_b200.so; do
 echo $lib; FTK_LIB_PATH=$PWD/$lib timeout 300 python tools/klt_ab.py affine direct 6 480 752 100 2000 2>&1 | tail -1 | cut -c1-200
done; done
timeout 600 python tools/fuzz_parity.py 1500 602 2>&1 | tail -1
