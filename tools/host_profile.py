"""cProfile of the host side of Simulation.step() for a bench config (where does the host time of a step go?):
    python tools/host_profile.py C2w 56"""
import cProfile
import os
import pstats
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

cfg_name, steps = sys.argv[1], int(sys.argv[2])
sim = bench.build_b200_sim(bench.CONFIGS[cfg_name], 1, fused=True, sort_period=4)
sim.step(30, keep_on_gpu=True)
pr = cProfile.Profile()
pr.enable()
sim.step(steps, keep_on_gpu=True)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats('cumulative').print_stats(28)
st.sort_stats('tottime').print_stats(14)
