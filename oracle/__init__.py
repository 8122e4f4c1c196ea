"""CPU checkers for voroffset_b200 - TEST INFRASTRUCTURE ONLY.

`oracle.cpu.Oracle`  : liboracle.so, our plain-C restatement (oracle.c).
`oracle.cpu.Reference`: oracle/_ref/libvoroffset_ref.so, the reference's own sources compiled in
place (see Makefile / ref_driver.cpp); present only where /root/reference was available at build
time (the .so then travels with the repo snapshot).

Nothing under voroffset_b200/ may import this package. Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs do.
"""
