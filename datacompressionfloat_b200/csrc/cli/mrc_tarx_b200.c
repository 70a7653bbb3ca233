/*
 * mrc_tarx_b200 -- multi-file front end with the reference's flags (src/main/mrc_tarx.c:345-420):
 *     mrc_tarx_b200 -i <file list> -o <output dir> [-t zip|unzip] [-b bits] [-n threads] [-d 0|1]
 * N worker threads like the reference's pool (mrc_tarx.c:41-176, queue adapt.c:337-356); every worker owns a GPU
 * context (on its own GPU when there are several), and small files are gathered into shared passes of the kernels.  Output names follow
 * adapt.c:297-304: x.mrc -> DIR/x.mrc.zip, x.mrc.zip -> DIR/x.mrc.  -d 1 sets isTestThroughput (no writes).
 */
#include <getopt.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../../include/mrczip_b200.h"

typedef struct {
    char **srcs, **dsts;
    int n, next;
    pthread_mutex_t lock;
    int unzip, bits, group;
    ctx_t total;
    int failed;
} job_t;

static const char *base_name(const char *p)
{
    const char *s = strrchr(p, '/');
    return s ? s + 1 : p;
}

static int ends_with(const char *s, const char *suf)
{
    const size_t a = strlen(s), b = strlen(suf);
    return a >= b && strcmp(s + a - b, suf) == 0;
}

static int load_list(job_t *j, const char *list, const char *dir)
{
    FILE *f = fopen(list, "r");
    if (!f) { fprintf(stderr, "cannot open list %s\n", list); return -1; }
    char line[4096];
    int cap = 16;
    j->srcs = malloc(sizeof(char *) * cap);
    j->dsts = malloc(sizeof(char *) * cap);
    while (fgets(line, sizeof line, f)) {
        size_t len = strlen(line);
        while (len && (line[len - 1] == '\n' || line[len - 1] == '\r' || line[len - 1] == ' ')) line[--len] = 0;
        if (!len) continue;
        if (j->n == cap) { cap *= 2; j->srcs = realloc(j->srcs, sizeof(char *) * cap); j->dsts = realloc(j->dsts, sizeof(char *) * cap); }
        char dst[4096];
        const char *bn = base_name(line);
        if (ends_with(bn, ".zip")) snprintf(dst, sizeof dst, "%s/%.*s", dir, (int)(strlen(bn) - 4), bn);
        else if (ends_with(bn, ".mrc")) snprintf(dst, sizeof dst, "%s/%s.zip", dir, bn);
        else { fprintf(stderr, "[%s:%d] Error: Only file with suffix [mrc | zip] can be processed\n", __FILE__, __LINE__); fclose(f); return -1; }
        j->srcs[j->n] = strdup(line);
        j->dsts[j->n] = strdup(dst);
        j->n++;
    }
    fclose(f);
    return 0;
}

/* Workers take runs of GROUP consecutive list entries and hand them to zip_compress_many / zip_uncompress_many: small
 * files of a run share passes of the GPU kernels (at most 32 chunks per pass), large ones go through one by one --
 * and every worker thread sits on its own GPU when there are several (the library deals devices round robin). */
#define GROUP 32

static void *worker(void *arg)
{
    job_t *j = (job_t *)arg;
    ctx_t ctx;
    init_context(&ctx);
    for (;;) {
        pthread_mutex_lock(&j->lock);
        const int i0 = j->next;
        int cnt = j->n - i0 < j->group ? j->n - i0 : j->group;
        if (cnt < 0) cnt = 0;
        j->next += cnt;
        pthread_mutex_unlock(&j->lock);
        if (cnt == 0) break;
        const int rc = j->unzip ? zip_uncompress_many(&ctx, cnt, (const char *const *)j->srcs + i0, (const char *const *)j->dsts + i0)
                                : zip_compress_many(&ctx, cnt, (const char *const *)j->srcs + i0, (const char *const *)j->dsts + i0, j->bits);
        if (rc != 0) { pthread_mutex_lock(&j->lock); j->failed++; pthread_mutex_unlock(&j->lock); }
    }
    pthread_mutex_lock(&j->lock);
    update_context(&j->total, &ctx);
    pthread_mutex_unlock(&j->lock);
    return NULL;
}

int main(int argc, char *argv[])
{
    const char *list = NULL, *dir = ".", *op = "zip";
    int threads = 2, opt;
    job_t j;
    memset(&j, 0, sizeof j);
    while ((opt = getopt(argc, argv, "hi:o:b:t:n:d:s:")) != -1) {
        switch (opt) {
            case 'i': list = optarg; break;
            case 'o': dir = optarg; break;
            case 'b': j.bits = atoi(optarg); break;
            case 't': op = optarg; break;
            case 'n': threads = atoi(optarg); break;
            case 'd': isTestThroughput = atoi(optarg); break;
            case 's': break; /* parsed but unused by the reference as well (mrc_tarx.c:357,393) */
            default:
                printf("Usage: %s -i <file list> -o <output dir> [-t zip|unzip] [-b bits] [-n threads] [-d 0|1]\n", argv[0]);
                return opt == 'h' ? 0 : 1;
        }
    }
    if (!list) { fprintf(stderr, "missing -i <file list>\n"); return 1; }
    j.unzip = strcmp(op, "unzip") == 0;
    if (load_list(&j, list, dir) != 0) return 1;
    if (threads < 1) threads = 1;
    /* runs of list entries per worker turn: everything in GROUPs, but never fewer turns than threads */
    j.group = j.n / threads < GROUP ? (j.n / threads > 0 ? j.n / threads : 1) : GROUP;
    if (threads > j.n) threads = j.n > 0 ? j.n : 1;
    pthread_mutex_init(&j.lock, NULL);
    init_context(&j.total);
    const double t0 = now_sec();
    pthread_t *th = malloc(sizeof(pthread_t) * threads);
    for (int i = 0; i < threads; i++) pthread_create(&th[i], NULL, worker, &j);
    for (int i = 0; i < threads; i++) pthread_join(th[i], NULL);
    const double dt = now_sec() - t0;
    const double bytes = (double)j.total.allFileSize;
    /* same line as mrc_tarx.c:231 */
    printf("num:%0.4f GBytes, time:%0.2f seconds, %0.2fMB/s\n", bytes / (1024.0 * 1024.0 * 1024.0), dt, bytes / (dt * 1024.0 * 1024.0));
    return j.failed ? 1 : 0;
}
