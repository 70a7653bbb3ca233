/*
 * mrc_tar_b200 -- single-file front end with the reference's flags (src/main/mrc_tar.c:104):
 *     mrc_tar_b200 -i <input> -o <output> [-t zip|unzip] [-b <bits to erase, 0..32>] [-s float]
 * Thin C: everything goes through run_compress / run_uncompress of libmrczip_b200.so.
 */
#include <getopt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../../include/mrczip_b200.h"

static void usage(const char *a0)
{
    printf("\nUsage:\n\n\t%s -i <input file> -o <output file> [-t <zip | unzip> -b <bits to erase>]\n", a0);
    printf("\t-b\tbits to be erased, range[0..32], default is 0\n");
    printf("\t-s\tdata type, only \"float\" is supported by the B200 path\n");
    printf("\t-t\toperation type [zip | unzip], default is zip\n\n");
}

int main(int argc, char *argv[])
{
    const char *in = NULL, *out = NULL, *op = "zip", *type = "float";
    int bits = 0, opt;
    if (argc < 2) { usage(argv[0]); return 1; }
    while ((opt = getopt(argc, argv, "hi:o:b:t:s:")) != -1) {
        switch (opt) {
            case 'i': in = optarg; break;
            case 'o': out = optarg; break;
            case 'b': bits = atoi(optarg); break;
            case 't': op = optarg; break;
            case 's': type = optarg; break;
            case 'h': usage(argv[0]); return 0;
            default: usage(argv[0]); return 1;
        }
    }
    if (!in || !out) { usage(argv[0]); return 1; }
    printf("%s\n", mzb_version());
    FILE *fin = fopen(in, "rb");
    if (!fin) { fprintf(stderr, "Error: [%s:%d]: Failed to  open input file :%s\n", __FILE__, __LINE__, in); return 1; }
    FILE *fout = fopen(out, "wb");
    if (!fout) { fprintf(stderr, "Error: [%s:%d]: Failed to open output file [%s] to write\n", __FILE__, __LINE__, out); return 1; }
    ctx_t ctx;
    init_context(&ctx);
    ctx.fileCount += 1;
    int rc;
    if (strcmp(op, "zip") == 0) {
        ctx.allFileSize += get_file_size(fin);
        rc = run_compress(fin, &ctx, fout, bits, type);
        print_context_info(&ctx, "Contex Info after Compression");
    } else if (strcmp(op, "unzip") == 0) {
        mrczip_header_t hd;
        init_mrczip_header(&hd, 0);
        if (read_mrczip_header(fin, &hd) != 0) { fclose(fin); fclose(fout); return 1; }
        print_mrczip_header(&hd, "Header Info in Decompression");
        rc = run_uncompress(fin, &ctx, &hd, fout, type);
        print_context_info(&ctx, "Contex Info after Decompression");
    } else {
        usage(argv[0]);
        rc = 1;
    }
    fclose(fin);
    fclose(fout);
    return rc == 0 ? 0 : 1;
}
