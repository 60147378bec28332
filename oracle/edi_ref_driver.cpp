/*
 * edi_ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Assembles EDI AF packets with the reference's OWN packetiser classes, compiled unmodified from
 * /root/reference/contrib/edioutput/{TagItems,TagPacket,AFPacket}.cpp and contrib/crc.c by oracle/Makefile.
 * Only the few lines of Output::EDI::write_frame around them (src/Outputs.cpp:194-263: time keeping and the order
 * of the TAG items) are restated here, because src/Outputs.cpp itself needs libzmq and the socket layer; the TAI
 * offset the reference asks ClockTAI for is an argument.
 *
 * usage: edi_ref_driver TIST DELAY_MS ALIGNMENT TAI_OFFSET START_TIME VERSION_TAG FRAME_LEN IN.bin OUT.bin [FEC [CHUNK_LEN]]
 *   IN.bin   records of FRAME_LEN bytes + 2 x int16 (peak left, right; host byte order)
 *   OUT.bin  per frame: uint32 packet size (host byte order) + the AF packet
 *   FEC      given (>= 0): the AF packet also goes through the reference's PFT layer (contrib/edioutput/PFT.cpp with
 *            contrib/ReedSolomon.cpp and contrib/fec, as Sender::write does for UDP: Transport.cpp:134-139) and
 *            OUT.bin holds, after each AF packet, uint32 fragment count and per fragment uint32 size + bytes
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "AFPacket.h"
#include "Log.h"
#include "PFT.h"
#include "TagItems.h"
#include "TagPacket.h"

Logger etiLog; /* the reference defines it in contrib/Globals.cpp, next to the remote control this driver has no use for */

int main(int argc, char **argv)
{
    if (argc < 10) return 2;
    const bool tist = atoi(argv[1]) != 0;
    const uint32_t delay_ms = (uint32_t)atol(argv[2]);
    const unsigned alignment = (unsigned)atoi(argv[3]);
    const int tai_offset = atoi(argv[4]);
    const std::time_t start_time = (std::time_t)atoll(argv[5]);
    const std::string version_tag = argv[6];
    const size_t frame_len = (size_t)atol(argv[7]);
    FILE *fi = fopen(argv[8], "rb"), *fo = fopen(argv[9], "wb");
    if (!fi || !fo) return 1;
    const int fec = argc > 10 ? atoi(argv[10]) : -1;
    edi::configuration_t conf;
    if (fec >= 0) conf.fec = (unsigned)fec;
    if (argc > 11) conf.chunk_len = (unsigned)atoi(argv[11]);
    edi::PFT pft(conf);

    edi::AFPacketiser afp;
    edi::TagDSTI tagDSTI;
    std::time_t edi_time = 0, send_version_at_time = 0;
    uint32_t timestamp = 0, num_seconds_sent = 0;
    std::vector<uint8_t> rec(frame_len + 4);
    while (fread(rec.data(), 1, rec.size(), fi) == rec.size()) {
        int16_t left, right;
        memcpy(&left, rec.data() + frame_len, 2);
        memcpy(&right, rec.data() + frame_len + 2, 2);
        // ---- src/Outputs.cpp:200-213
        if (edi_time == 0) {
            edi_time = start_time + (delay_ms / 1000);
            send_version_at_time = edi_time;
            for (int32_t sub_ms = (delay_ms % 1000); sub_ms > 0; sub_ms -= 24) timestamp += 24 << 14;
        }
        // ---- src/Outputs.cpp:215-258
        edi::TagStarPTR tagStarPtr("DSTI");
        tagDSTI.stihf = false;
        tagDSTI.atstf = tist;
        timestamp += 24 << 14;
        if (timestamp > 0xf9FFff) {
            timestamp -= 0xfa0000;
            edi_time += 1;
            num_seconds_sent++;
        }
        tagDSTI.set_edi_time(edi_time, tai_offset);
        tagDSTI.tsta = timestamp & 0xffffff;
        tagDSTI.rfadf = false;
        edi::TagSSm tagPayload;
        tagPayload.istd_data = rec.data();
        tagPayload.istd_length = frame_len;
        edi::TagODRAudioLevels tagAudioLevels(left, right);
        edi::TagODRVersion tagVersion(version_tag, num_seconds_sent);
        edi::TagPacket tagpacket(alignment);
        tagpacket.tag_items.push_back(&tagStarPtr);
        tagpacket.tag_items.push_back(&tagDSTI);
        tagpacket.tag_items.push_back(&tagPayload);
        tagpacket.tag_items.push_back(&tagAudioLevels);
        if (send_version_at_time < edi_time) {
            send_version_at_time += 10;
            tagpacket.tag_items.push_back(&tagVersion);
        }
        // ---- contrib/edioutput/Transport.cpp:126-132
        const edi::AFPacket af = afp.Assemble(tagpacket);
        const uint32_t n = (uint32_t)af.size();
        fwrite(&n, 4, 1, fo);
        fwrite(af.data(), 1, af.size(), fo);
        if (fec >= 0) {
            const std::vector<edi::PFTFragment> frags = pft.Assemble(af);
            const uint32_t nf = (uint32_t)frags.size();
            fwrite(&nf, 4, 1, fo);
            for (const auto &f : frags) {
                const uint32_t fs = (uint32_t)f.size();
                fwrite(&fs, 4, 1, fo);
                fwrite(f.data(), 1, f.size(), fo);
            }
        }
    }
    fclose(fi);
    fclose(fo);
    return 0;
}
