// Test harness for PyIntSet (pandelos_b200/csrc/host/girvan_newman.h): reads non-negative ints from stdin, adds them
// one at a time, prints the set's iteration order — tests/test_netclu_cc.py compares it with the running CPython's.
#include <cstdio>

#include "girvan_newman.h"

int main() {
    pd_host::PyIntSet set;
    unsigned long long v;
    while (scanf("%llu", &v) == 1) set.add(static_cast<uint32_t>(v));
    for (uint32_t k : set.order()) printf("%u\n", k);
    return 0;
}
