"""oracle/ — TEST INFRASTRUCTURE ONLY (CPU restatement of the reference's detection forward path).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package. The product (multipathnet_b200/) never does; it fails loudly without
libmpn_b200.so. See DESIGN.md "Oracle" for what is pinned against the reference and what is
"parity unpinned".
"""
