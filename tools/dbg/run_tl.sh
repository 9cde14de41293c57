timeout 200 python profiles/timeline.py elec > /dev/null 2>&1; sed -n 1,16p gpurun_out/timeline_elec.txt
