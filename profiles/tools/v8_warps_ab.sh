set -x
SAMPLES_BIG=296
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for np in 200 240; do for pn in 0 1; do
  echo "== NPTS $np PATCHNORM $pn new"; NPTS=$np SAMPLES=296 FRAMES=30 DONORM=$pn PATCHNORM=$pn timeout 300 python profiles/tools/run_chain.py | tail -1
  echo "== NPTS $np PATCHNORM $pn old"; ICT_FAST_V1=1 NPTS=$np SAMPLES=296 FRAMES=30 DONORM=$pn PATCHNORM=$pn timeout 300 python profiles/tools/run_chain.py | tail -1
done; done
echo "== NPTS 100 8 warps 1 sample";  NPTS=100 SAMPLES=1 FRAMES=100 timeout 300 python profiles/tools/run_chain.py | tail -1
echo "== NPTS 100 16 warps 1 sample"; ICT_V8_WARPS=16 NPTS=100 SAMPLES=1 FRAMES=100 timeout 300 python profiles/tools/run_chain.py | tail -1
echo "== NPTS 100 8 warps 592";  NPTS=100 SAMPLES=592 FRAMES=30 timeout 300 python profiles/tools/run_chain.py | tail -1
echo "== NPTS 100 16 warps 592"; ICT_V8_WARPS=16 NPTS=100 SAMPLES=592 FRAMES=30 timeout 300 python profiles/tools/run_chain.py | tail -1
