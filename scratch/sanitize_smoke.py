"""One small fwd+bwd of the production kernels (C = 256: tcgen05 K1/K3/K4/K6, cluster Cholesky, small-matrix chain) for
compute-sanitizer:  compute-sanitizer --tool {memcheck,racecheck,synccheck,initcheck} python scratch/sanitize_smoke.py"""
import sys
sys.path.insert(0, '/root/repo')
import __graft_entry__ as G
G.smoke()
