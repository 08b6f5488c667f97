#!/bin/bash
# dev tool: phases of the command-line driver on the C3 mesh read from an OBJ file (one B200)
python - <<'P'
import os, sys
sys.path.insert(0, os.getcwd())
import __graft_entry__ as entry
api = entry.load_package().api
api.write_obj("/tmp/rtb_hf708.obj", api.heightfield_mesh(708, 20 * 1920 / 1080 * 0.98))
P
TIMEFORMAT="wall %R s"
for i in 1 2; do
  time (RTB_TIMING=1 raytracer.c_b200/bin/raytracer -w 1920 -h 1080 -s 128 -o /tmp/rtb_c3.png \
      -c obj:/tmp/rtb_hf708.obj -m 1,1,1,0,0,0,0 2>&1 | grep -v "^\[" | grep "main:\|took\|load_obj:\|render:\|rays\|done")
done
rm -f /tmp/rtb_hf708.obj /tmp/rtb_c3.png
