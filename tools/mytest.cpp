// mytest.cpp -- Test.cpp-shaped CLI over the drop-in Coder (include/MyLdpc.h):
//   mytest <srcLength bytes> <batchSize> <snr dB> <SP|MS|CPU|TDMP|TDMPCL|MSCL> [seed] [gpus]
// Same flow as the reference's harness (Test.cpp:15-118): payload 'a'..'z' -> encode -> BPSK + AWGN
// with sd = 10^(-snr/20) -> decode -> byte compare; prints the same lines (sd=, <alg>:<seconds>,
// ErrNum=, ThroughPut= in payload bytes/s) plus the mean iteration count.  Every algorithm name
// runs the CUDA min-sum decoder.  Optional seed makes the run repeatable (the reference seeds
// with time(0)); optional gpus shards the codewords over that many devices.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <vector>

#include "MyLdpc.h"

int main(int argc, char **argv) {
    if (argc < 5) {
        std::fprintf(stderr, "usage: %s srcLength batchSize snr_dB SP|MS|CPU|TDMP|TDMPCL|MSCL [seed] [shards] [visible_gpus]\n", argv[0]);
        return 2;
    }
    const int z = 24, ldpcN = z * 24, ldpcK = ldpcN / 4 * 3;   // Test.cpp:19-26
    Coder coder(ldpcK, ldpcN, rate_3_4_b);
    srand(argc > 5 ? (unsigned)atoi(argv[5]) : (unsigned)time(0));
    const int gpus = argc > 6 ? atoi(argv[6]) : 1;
    const int visible = argc > 7 ? atoi(argv[7]) : gpus;  // shards wrap around the visible devices
    if (gpus > 1) {
        std::vector<int> ids(gpus);
        for (int i = 0; i < gpus; ++i) ids[i] = i % (visible > 0 ? visible : 1);
        coder.setDevices(ids.data(), gpus);
    }
    const int srcLength = atoi(argv[1]);
    std::vector<char> srcCode(srcLength), priorCode(coder.getPriorCodeLength(srcLength)), newSrcCode(srcLength + 1);
    std::vector<float> postCode(coder.getPostCodeLength(srcLength));
    for (int i = 0; i < srcLength; i++) srcCode[i] = 'a' + i % 26;
    if (coder.forEncoder() || coder.forDecoder(atoi(argv[2]))) {
        std::fprintf(stderr, "setup failed: %s\n", coder.lastError());
        return 1;
    }
    coder.encode(srcCode.data(), priorCode.data(), srcLength);
    const float snr = (float)atof(argv[3]);
    const float sd = 1 / (pow(10, snr / 20));
    std::cout << "sd=" << sd << std::endl;
    coder.test(priorCode.data(), postCode.data(), coder.getPriorCodeLength(srcLength), sd);

    decodeType t = DecodeMS;
    const char *names[] = {"CPU", "MS", "SP", "TDMP", "TDMPCL", "MSCL"};
    for (int i = 0; i < 6; ++i)
        if (!strcmp(argv[4], names[i])) t = (decodeType)i;
    if (coder.addDecodeType(t)) {
        std::fprintf(stderr, "addDecodeType failed: %s\n", coder.lastError());
        return 1;
    }
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = coder.decode(postCode.data(), newSrcCode.data(), srcLength, t);
    const double decodeTime = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (rc) {
        std::fprintf(stderr, "decode failed: %s\n", coder.lastError());
        return 1;
    }
    std::cout << argv[4] << ":" << decodeTime << std::endl;
    int errNum = 0;
    for (int i = 0; i < srcLength; ++i)
        if (srcCode[i] != newSrcCode[i]) ++errNum;
    double it = 0;
    for (int i = 0; i < coder.lastCodeSize(); ++i) it += coder.lastIterations()[i];
    std::cout << "ErrNum=" << errNum << std::endl;
    std::cout << "ThroughPut=" << srcLength / decodeTime << std::endl;
    std::cout << "MeanIterations=" << (coder.lastCodeSize() ? it / coder.lastCodeSize() : 0.0) << std::endl;
    return 0;
}
