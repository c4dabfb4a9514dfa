import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.probe_par import run
for c in [int(a) for a in sys.argv[1:]] or [4]:
    run(16, 512, c, 20, 0)
