// The calls a QB3 user makes that need no device: handle life cycle, setters, size bound, header parsing.
#include <QB3.h>
#include <cstdio>
#include <cstring>
int main()
{
    encsp e = qb3_create_encoder(512, 512, 3, QB3_U8);
    if (!e) return 1;
    size_t map[QB3_MAXBANDS] = {1, 1, 1};
    if (!qb3_set_encoder_coreband(e, 3, map) || qb3_set_encoder_mode(e, QB3M_BEST) != QB3M_BEST) return 2;
    const size_t bound = qb3_max_encoded_size(e);
    qb3_destroy_encoder(e);
    if (bound != 1024 + (size_t)((17.0 / 16.0 + 8) * (16.0 * 128 * 128 * 3) / 8)) return 3;
    const unsigned char k4[] = {0x51, 0x42, 0x33, 0x80, 7, 0, 7, 0, 0, 0, 8, 0x53, 0x43, 8, 0, 0x23, 0x76, 0xfb, 0xae,
                                0xd9, 0x8c, 0x54, 0x01, 0x44, 0x54, 0};
    size_t sz[3];
    decsp d = qb3_read_start((void *)k4, sizeof(k4), sz);
    if (!d || sz[0] != 8 || sz[1] != 8 || sz[2] != 1 || !qb3_read_info(d) || qb3_get_mode(d) != QB3M_FTL) return 4;
    qb3_destroy_decoder(d);
    std::puts("QB3 package ok");
    return 0;
}
