#ifndef STUB_NUMAIF_H
#define STUB_NUMAIF_H
#define _GNU_SOURCE 1
#include <unistd.h>
#include <sys/syscall.h>
#define MPOL_F_NODE (1<<0)
#define MPOL_F_ADDR (1<<1)
static inline long get_mempolicy(int *mode, unsigned long *nodemask, unsigned long maxnode, void *addr, unsigned long flags)
{ return syscall(SYS_get_mempolicy, mode, nodemask, maxnode, addr, flags); }
#endif
