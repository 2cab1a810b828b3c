"""TEST INFRASTRUCTURE ONLY: CPU checkers for the CUDA LLD path (see oracle/osm_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (opensmile_b200/) never does.
"""
