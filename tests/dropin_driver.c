/*
 * dropin_driver.c -- TEST INFRASTRUCTURE: the call sequence of odr-audioenc for --dab
 * (src/odr-audioenc.cpp:687-721 set-up, :1139-1158 per frame, :1161 finish) written against <toolame.h> only.
 * It is compiled twice by tests/test_dropin_link.py: with the reference's own libtoolame-dab/toolame.h and with
 * include/toolame.h, and linked against libtoolame_b200.so both times -- the drop-in claim at the C level.
 * usage: dropin_driver FS MODE BITRATE PSY IN.pcm OUT.mp2
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <toolame.h>

int main(int argc, char **argv)
{
    if (argc < 7) return 2;
    long fs = atol(argv[1]);
    char mode = argv[2][0];
    int nch = mode == 'm' ? 1 : 2;
    FILE *fi = fopen(argv[5], "rb"), *fo = fopen(argv[6], "wb");
    if (!fi || !fo) return 1;
    if (toolame_init() || toolame_set_samplerate(fs) || toolame_set_psy_model(atoi(argv[4])) ||
        toolame_set_channel_mode(mode) || toolame_set_bitrate(atoi(argv[3])) || toolame_set_pad(0))
        return 3;
    static short inter[2 * 1152], buf[2][1152];
    static unsigned char out[4092];
    while (fread(inter, 2, (size_t)nch * 1152, fi) == (size_t)nch * 1152) {
        for (int i = 0; i < 1152; i++)
            for (int ch = 0; ch < nch; ch++) buf[ch][i] = inter[i * nch + ch];
        int n = toolame_encode_frame(buf, NULL, 0, out, sizeof out);
        if (n > 0) fwrite(out, 1, (size_t)n, fo);
    }
    int n = toolame_finish(out, sizeof out);
    if (n > 0) fwrite(out, 1, (size_t)n, fo);
    fclose(fi);
    fclose(fo);
    return 0;
}
