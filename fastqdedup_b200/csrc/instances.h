// instances.h -- the (K, PW) instantiations of this build: key length <= 32*PW symbols, alphabet
// (+PAD) < 2^K.  Grouped so that pipeline_inst.cu compiles as several translation units in parallel
// (-DFQD_GROUP=n).  The reference has no key-length limit below 2^32 (TRIE_NODE_SUFFIX_MAX_SIZE,
// _triemodule.c:115); this build packs keys of up to 640 symbols for alphabets of up to 15 symbols (DNA with
// every IUPAC code) and 320 symbols beyond that, which covers un-sliced 2 x 300 nt read pairs (the CLI
// default without --check-lengths, __init__.py:232, :251).  Longer keys: FQD_ERR_UNSUPPORTED.
#pragma once

#define FQD_INSTANCES_G0(X) X(3, 1) X(3, 2)
#define FQD_INSTANCES_G1(X) X(3, 3) X(3, 4) X(3, 5)
#define FQD_INSTANCES_G2(X) X(3, 8) X(3, 10) X(3, 16) X(3, 20)
#define FQD_INSTANCES_G3(X) X(4, 1) X(4, 2) X(4, 4) X(4, 8)
#define FQD_INSTANCES_G4(X) X(4, 10) X(4, 16) X(4, 20)
#define FQD_INSTANCES_G5(X) X(8, 1) X(8, 2) X(8, 4)
#define FQD_INSTANCES_G6(X) X(8, 8) X(8, 10)
#define FQD_N_GROUPS 7

#define FQD_INSTANCES(X)                                                                  \
    FQD_INSTANCES_G0(X) FQD_INSTANCES_G1(X) FQD_INSTANCES_G2(X) FQD_INSTANCES_G3(X)       \
    FQD_INSTANCES_G4(X) FQD_INSTANCES_G5(X) FQD_INSTANCES_G6(X)
