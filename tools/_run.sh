python tools/_diag.py > gpurun_out/t9_diag.txt 2>&1
LPSR_NO_NSPLIT=1 python tools/_diag.py > gpurun_out/t9_diag_nonsplit.txt 2>&1
