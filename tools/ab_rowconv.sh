# per-launch times of the row-streaming kernel inside the forward under the LPSR_UMMA_DEBUG profiling switches (1 no MMAs, 2 no stores, 4 no TMA loads)
for dbg in 0 1 2 4 3 5 6 7; do
LPSR_UMMA_DEBUG=$dbg python tools/layer_times.py --filter rowconv 2>&1 | tail -1
done
LPSR_ROWCONV=0 python tools/layer_times.py --filter rd 2>&1 | tail -1
