// bbfft_tools.cpp -- command line tools of the CUDA backend, one binary with three personalities
// (selected by argv[0] or the first argument):
//   bbfft-aot-generate [-d <arch>] [-i <device info>] <kernel file> <descriptor>...
//       FFT descriptors -> cubin (NVRTC, no GPU needed).  Role of the reference's
//       tools/aot/main.cpp:22-60 (ocloc -> native binary); load the blob with
//       bbfft::cuda::create_aot_module and register it in a bbfft::aot_cache.
//   bbfft-offline-generate [-i <device info>] <descriptor>...
//       prints the CUDA C++ kernel stubs.  Role of tools/offline/main.cpp:15-33.
//   bbfft-device-info [<device ordinal>]
//       prints the device_info string of a CUDA device.  Role of tools/device_info/main.cpp:10-15.
#include "bbfft/cuda/device.hpp"
#include "bbfft/cuda/online_compiler.hpp"
#include "bbfft/generator.hpp"
#include "bbfft/parser.hpp"

#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

using namespace bbfft;

namespace {
// built-in device table (reference: tools/common/info.cpp:8-9 has only "pvc")
const std::map<std::string, std::pair<std::string, device_info>> builtin = {
    {"b200", {"sm_100a", device_info{1024, {32}, 232448, device_type::gpu}}},
    {"sm_100a", {"sm_100a", device_info{1024, {32}, 232448, device_type::gpu}}},
};

void help(std::ostream &os, std::string const &tool) {
    if (tool == "bbfft-aot-generate") {
        os << "usage: bbfft-aot-generate [-d device] [-i device_info] kernel_file descriptor...\n"
              "  -d, --device       target: b200 | sm_100a (default b200)\n"
              "  -i, --device_info  \"{max_wg, {32}, smem_bytes, gpu}\" (overrides the built-in table)\n"
              "  kernel_file        output cubin\n"
              "  descriptor         e.g. scfo16.64*1000, drfo64x64x64*8 (docs: descriptor format)\n";
    } else if (tool == "bbfft-offline-generate") {
        os << "usage: bbfft-offline-generate [-i device_info] descriptor...\n";
    } else {
        os << "usage: bbfft-device-info [device ordinal]\n";
    }
}
} // namespace

int main(int argc, char **argv) {
    std::string tool = argv[0];
    auto slash = tool.find_last_of('/');
    if (slash != std::string::npos) tool = tool.substr(slash + 1);
    int first = 1;
    if (tool.rfind("bbfft-", 0) != 0 && argc > 1) {
        tool = argv[1];
        first = 2;
    }
    try {
        if (tool == "bbfft-device-info") {
            int dev = argc > first ? std::atoi(argv[first]) : 0;
            std::cout << get_device_info(dev) << std::endl;
            return 0;
        }
        std::string device = "b200", out_file;
        device_info info = {};
        std::vector<configuration> cfgs;
        const bool aot = tool == "bbfft-aot-generate";
        for (int i = first; i < argc; ++i) {
            std::string a = argv[i];
            if (a == "-h" || a == "--help") {
                help(std::cout, tool);
                return 0;
            } else if ((a == "-d" || a == "--device") && i + 1 < argc) {
                device = argv[++i];
            } else if ((a == "-i" || a == "--device_info") && i + 1 < argc) {
                info = parse_device_info(argv[++i]);
            } else if ((a == "-f" || a == "--format") && i + 1 < argc) {
                if (std::strcmp(argv[++i], "native") != 0) {
                    throw std::runtime_error("==> Error: the CUDA backend produces native modules only");
                }
            } else if (a[0] == '-') {
                throw std::runtime_error("==> Error: unrecognized argument " + a);
            } else if (aot && out_file.empty()) {
                out_file = a;
            } else {
                cfgs.emplace_back(parse_fft_descriptor(a));
            }
        }
        if (aot && out_file.empty()) throw std::invalid_argument("==> You need to provide the kernel file name");
        if (cfgs.empty()) throw std::invalid_argument("==> You need to provide at least one FFT desciptor");
        std::string arch = "sm_100a";
        auto it = builtin.find(device);
        if (it != builtin.end()) {
            arch = it->second.first;
            if (info.max_work_group_size == 0) info = it->second.second;
        } else if (info.max_work_group_size == 0) {
            throw std::invalid_argument("Device info missing for device \"" + device +
                                        "\". You need to provide device info via --device_info.");
        } else {
            arch = device;
        }
        std::ostringstream src;
        auto names = generate_fft_kernels(src, cfgs, info);
        if (!aot) {
            std::cout << src.str();
            return 0;
        }
        auto bin = cuda::compile_to_native(src.str(), arch);
        std::ofstream f(out_file, std::ios::binary);
        f.write(reinterpret_cast<char const *>(bin.data()), std::streamsize(bin.size()));
        if (!f) throw std::runtime_error("==> Error: cannot write " + out_file);
        for (auto const &n : names) std::cerr << n << std::endl;
        return 0;
    } catch (std::exception const &e) {
        std::cerr << e.what() << std::endl << std::endl;
        help(std::cerr, tool);
        return -1;
    }
}
