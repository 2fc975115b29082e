for c in 8 2 1; do NB200_ADAM_CTAS_PER_SM=$c python profiles/overlap_probe.py 2>&1 | grep -v Warning | tail -12; done > gpurun_out/r02h_overlap.txt 2>&1; cat gpurun_out/r02h_overlap.txt
