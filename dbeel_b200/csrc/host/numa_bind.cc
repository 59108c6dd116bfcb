// numa_bind.cc -- keep a shard's host thread and its pinned buffers next to its GPU.
//
// dbeel runs one executor thread per core and pins it there (src/main.rs:51-60, glommio LocalExecutorBuilder with
// Placement::Fixed(cpu)).  With one GPU per shard the host side of a compaction is two PCIe streams through the root
// complex of ONE socket: staging buffers that live on the other socket cross the inter-socket link twice and the 8-GPU
// end-to-end rate collapses (measured in round 1: 0.52 efficiency at N = 8 with floating threads).  dbeel_bind_to_gpu moves
// the calling thread onto the CPUs of the GPU's NUMA node and makes that node the preferred one for its future page
// allocations, so memory pinned afterwards (dbeel_host_alloc, cudaHostAlloc) is first-touched locally.
//
// Topology comes from sysfs: /sys/bus/pci/devices/<bdf>/numa_node and /sys/devices/system/node/node<N>/cpulist.
#include <cuda_runtime.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <string>

#include "../../../include/dbeel_compact.h"

namespace {

bool read_line(const std::string &path, char *buf, size_t cap) {
    FILE *f = fopen(path.c_str(), "r");
    if (!f) return false;
    const bool ok = fgets(buf, (int)cap, f) != nullptr;
    fclose(f);
    return ok;
}

// "0-31,64-95" -> cpu_set_t
int parse_cpulist(const char *s, cpu_set_t *set) {
    CPU_ZERO(set);
    int n = 0;
    while (*s) {
        char *end;
        long a = strtol(s, &end, 10);
        if (end == s) break;
        long b = a;
        if (*end == '-') {
            s = end + 1;
            b = strtol(s, &end, 10);
        }
        for (long c = a; c <= b && c < CPU_SETSIZE; c++) { CPU_SET((int)c, set); n++; }
        s = *end == ',' ? end + 1 : end;
        if (*end != ',' ) break;
    }
    return n;
}

} // namespace

extern "C" int dbeel_gpu_numa_node(int device) {
    char bdf[32] = {0};
    if (cudaDeviceGetPCIBusId(bdf, sizeof bdf, device) != cudaSuccess) { cudaGetLastError(); return -1; }
    for (char *c = bdf; *c; c++) *c = (char)((*c >= 'A' && *c <= 'Z') ? *c + 32 : *c); // sysfs names are lower case
    char line[64];
    if (!read_line(std::string("/sys/bus/pci/devices/") + bdf + "/numa_node", line, sizeof line)) return -1;
    return atoi(line); // -1 on single-node machines
}

extern "C" int dbeel_bind_to_gpu(int device, int *numa_node, int *n_cpus) {
    if (numa_node) *numa_node = -1;
    if (n_cpus) *n_cpus = 0;
    const int node = dbeel_gpu_numa_node(device);
    if (node < 0) return DBEEL_OK; // no NUMA information: nothing to do, not an error
    char line[4096];
    if (!read_line("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist", line, sizeof line)) return DBEEL_OK;
    cpu_set_t want, have, both;
    if (parse_cpulist(line, &want) == 0) return DBEEL_OK;
    // stay inside what the process is allowed to use (cgroup cpusets, taskset)
    if (sched_getaffinity(0, sizeof have, &have) != 0) return DBEEL_ERR_INVALID_ARG;
    CPU_AND(&both, &want, &have);
    if (CPU_COUNT(&both) == 0) return DBEEL_OK;
    if (sched_setaffinity(0, sizeof both, &both) != 0) return DBEEL_ERR_INVALID_ARG;
    // set_mempolicy(MPOL_PREFERRED, {node}): future pages of this thread come from the GPU's node when it has room
    unsigned long mask[16] = {0};
    if (node < (int)(sizeof mask * 8)) {
        mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
#ifdef SYS_set_mempolicy
        (void)syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, mask, sizeof mask * 8);
#endif
    }
    if (numa_node) *numa_node = node;
    if (n_cpus) *n_cpus = CPU_COUNT(&both);
    return DBEEL_OK;
}
