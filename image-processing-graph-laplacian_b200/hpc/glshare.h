/* glshare.h -- hand-over of a buffer that every rank (forked process) fills a part of and rank 0 reads, used repeatedly:
 * the u8 result image (ComputeResultFromLaplacian) and every eigenvector column dump (WriteMatCol / WritePngMatCol) go
 * through the same shared mapping.  Replaces the gather-to-rank-0 of the reference (VecScatterCreateToZero,
 * hpc/utils.c:498-506).  Two monotonic counters in shared memory, no reset, so a fast rank can never run into the next
 * use while rank 0 still reads the previous one:
 *     done     += 1 by every rank when its part of generation g is written      (rank 0 waits for (g + 1) * size)
 *     consumed  = g + 1 by rank 0 when it has read generation g                 (the others wait for it before writing g + 1)
 * Header-only and free of CUDA so that tests/test_host.py can exercise it with plain forked processes. */
#ifndef GLB200_GLSHARE_H
#define GLB200_GLSHARE_H
#include <sys/types.h>
#include <unistd.h>

typedef struct GLShare {
    volatile int* done;      /* shared */
    volatile int* consumed;  /* shared */
    int rank, size;
    int gen;                 /* uses this process has completed */
    pid_t parent;            /* ranks > 0: pid of rank 0's process (0 = not checked); a rank whose parent is gone gives up */
} GLShare;

/* every rank, before it writes its part */
static inline void GLShareBegin(GLShare* s)
{
    while (*s->consumed < s->gen) {
        /* rank 0 died (it alone advances `consumed`): nobody will ever read this generation */
        if (s->rank != 0 && s->parent > 0 && getppid() != s->parent) _exit(1);
        usleep(100);
    }
    __sync_synchronize();
}

/* every rank, after it has written its part */
static inline void GLShareDone(GLShare* s)
{
    __sync_synchronize();
    __sync_fetch_and_add((int*)s->done, 1);
}

/* rank 0: wait for every part of this generation; `alive` (may be NULL) is polled so that a dead rank does not hang the
 * wait: return non-zero from it to give up.  Returns 0 when all parts are there. */
static inline int GLShareWait(GLShare* s, int (*alive)(void*), void* arg)
{
    while (*s->done < (s->gen + 1) * s->size) {
        if (alive && alive(arg)) return 1;
        usleep(200);
    }
    __sync_synchronize();
    return 0;
}

/* every rank, when it is done with this generation (rank 0: after reading the buffer) */
static inline void GLShareRelease(GLShare* s)
{
    if (s->rank == 0) {
        __sync_synchronize();
        *s->consumed = s->gen + 1;
    }
    s->gen++;
}
#endif
