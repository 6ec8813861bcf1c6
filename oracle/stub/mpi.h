/* ORACLE SCAFFOLDING — NOT PRODUCT CODE.
 *
 * Single-process stand-in for <mpi.h>, used only to compile the UNMODIFIED reference
 * (/root/reference/*.cpp) into oracle/_ref/main_ref in a container that has no MPI.
 * The reference's ranks never exchange payload: the only data that crosses ranks is the
 * record-file path (main.cpp:27,35); everything else is rank/size queries, barriers and a timer.
 *
 *   rank / size     <- env ZWZ_STUB_RANK / ZWZ_STUB_SIZE (default 0 / 1)
 *   MPI_Bcast       <- rank 0 writes $ZWZ_STUB_DIR/bcast_<k>, other ranks poll-read it
 *   MPI_Barrier     <- no-op (emulated ranks are started rank-0-first by the test harness)
 *   MPI_Wtime       <- CLOCK_MONOTONIC
 */
#ifndef ZWZ_ORACLE_STUB_MPI_H
#define ZWZ_ORACLE_STUB_MPI_H
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;
enum { MPI_COMM_WORLD = 0 };
enum { MPI_CHAR = 1, MPI_UNSIGNED_LONG_LONG = 8 }; /* value == element size in bytes */

static inline int zwz_stub_env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}
static inline int MPI_Init(int *, char ***) { return 0; }
static inline int MPI_Finalize(void) { return 0; }
static inline int MPI_Comm_rank(MPI_Comm, int *rank) { *rank = zwz_stub_env_int("ZWZ_STUB_RANK", 0); return 0; }
static inline int MPI_Comm_size(MPI_Comm, int *size) { *size = zwz_stub_env_int("ZWZ_STUB_SIZE", 1); return 0; }
static inline int MPI_Barrier(MPI_Comm) { return 0; }
static inline int MPI_Abort(MPI_Comm, int code) { exit(code); return 0; }
static inline double MPI_Wtime(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}
static inline int MPI_Bcast(void *buf, int count, MPI_Datatype type, int /*root*/, MPI_Comm) {
    static int seq = 0;
    const char *dir = getenv("ZWZ_STUB_DIR");
    int k = seq++;
    if (!dir || zwz_stub_env_int("ZWZ_STUB_SIZE", 1) <= 1) return 0; /* one rank: nothing to exchange */
    size_t bytes = (size_t) count * (size_t) type;
    char path[4096];
    snprintf(path, sizeof path, "%s/bcast_%d", dir, k);
    if (zwz_stub_env_int("ZWZ_STUB_RANK", 0) == 0) {
        char tmp[4200];
        snprintf(tmp, sizeof tmp, "%s.tmp", path);
        FILE *f = fopen(tmp, "wb");
        if (!f) return 1;
        fwrite(buf, 1, bytes, f);
        fclose(f);
        rename(tmp, path); /* atomic publish so a polling reader never sees a partial file */
    } else {
        for (int tries = 0; tries < 60000; ++tries) { /* up to ~60 s */
            FILE *f = fopen(path, "rb");
            if (f) {
                size_t got = fread(buf, 1, bytes, f);
                fclose(f);
                return got == bytes ? 0 : 1;
            }
            usleep(1000);
        }
        return 1;
    }
    return 0;
}
#endif
