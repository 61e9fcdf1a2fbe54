/* CPU harness for hpc/glshare.h (tests/test_host.py): `size` forked processes hand a shared buffer over `gens` times.
 * Every rank writes its slot of generation g, rank 0 checks that it sees generation g in EVERY slot (never an older or a
 * newer one: a rank must not start writing g + 1 before rank 0 has read g), with random delays on both sides. */
#include <stdio.h>
#include <stdlib.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include "glshare.h"

int main(int argc, char** argv)
{
    const int size = argc > 1 ? atoi(argv[1]) : 3, gens = argc > 2 ? atoi(argv[2]) : 300;
    unsigned char* page = (unsigned char*)mmap(NULL, 4096, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (page == MAP_FAILED) return 2;
    GLShare s = {(volatile int*)page, (volatile int*)(page + 256), 0, size, 0};
    volatile int* slot = (volatile int*)(page + 512);
    pid_t kids[64] = {0};
    for (int r = 1; r < size; ++r) {
        pid_t pid = fork();
        if (pid == 0) { s.rank = r; break; }
        kids[r] = pid;
    }
    srand(1234u + (unsigned)s.rank);
    int bad = 0;
    for (int g = 0; g < gens; ++g) {
        GLShareBegin(&s);
        if (rand() % 3 == 0) usleep((useconds_t)(rand() % 300));
        slot[s.rank] = g * 1000 + s.rank;
        GLShareDone(&s);
        if (s.rank == 0) {
            if (GLShareWait(&s, NULL, NULL)) return 3;
            if (rand() % 3 == 0) usleep((useconds_t)(rand() % 300));      /* a slow reader */
            for (int r = 0; r < size; ++r)
                if (slot[r] != g * 1000 + r) {
                    fprintf(stderr, "generation %d: slot %d holds %d\n", g, r, slot[r]);
                    bad = 1;
                }
        }
        GLShareRelease(&s);
    }
    if (s.rank != 0) _exit(0);
    for (int r = 1; r < size; ++r) {
        int st = 0;
        waitpid(kids[r], &st, 0);
        if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) bad = 1;
    }
    printf("%s\n", bad ? "FAILED" : "ok");
    return bad;
}
