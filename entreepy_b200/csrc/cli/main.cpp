// entreepy — command-line driver over libentreepy_b200 (same grammar as the reference's
// src/main.zig:42-208: options before or after the command, `c FILE` / `d FILE`, -o PATH).
// It is the boundary caller of the hot path, not part of it: it reads the whole input into
// pinned memory, calls et_encode / et_decode once and writes the result with one write.
//
//   entreepy [-h] [-p] [-t] [-d] c|d FILE [-o OUT]
//
// Differences from the reference, all deliberate:
//   * the default output path works (main.zig:154-170 reads an undefined slice; the documented
//     behaviour `[file].et` / `decoded_[file]` is what this does);
//   * --strict (extension) checks the magic e7 c0 de and the version byte before decoding
//     (the reference skips file[0..4) unchecked, TODO at main.zig:199);
//   * errors are printed as "error: <name>" and the exit code is 1.
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../../include/entreepy_b200.h"

namespace {

const char kHelp[] =
    "Entreepy - Text compression tool (B200 build)\n\n"
    "Usage: entreepy [options] [command] [file] [command options]\n\n"
    "Options:\n"
    "    -h, --help     show help\n"
    "    -p, --print    print decompressed text to stdout\n"
    "    -t, --test     test/dry run, does not write to file\n"
    "    -d, --debug    print huffman code dictionary and performance times to stdout\n"
    "        --strict   refuse input that does not start with the .et magic and version\n\n"
    "Commands:\n"
    "    c    compress a file\n"
    "    d    decompress a file\n\n"
    "Command Options:\n"
    "    -o, --output    output file (default: [file].et or decoded_[file])\n\n"
    "Examples:\n"
    "    entreepy -d c text.txt -o text.txt.et\n"
    "    entreepy -ptd d text.txt.et -o decoded_text.txt\n";

enum class Mode { None, Compress, Decompress };

struct Options {
    bool print = false, debug = false, dry = false, strict = false, help = false;
    Mode mode = Mode::None;
    std::string in_path, out_path;
};

int fail(const char *what, const std::string &detail) {
    std::fprintf(stderr, "error: %s%s%s\n", what, detail.empty() ? "" : ": ", detail.c_str());
    return 1;
}

// main.zig:69-146
int parse(int argc, char **argv, Options *o) {
    enum { Normal, OutPath, InPath } state = Normal;
    if (argc <= 1) {
        o->help = true;
        return 0;
    }
    for (int i = 1; i < argc; ++i) {
        const std::string arg = argv[i];
        if (state == InPath) {
            o->in_path = arg;
            state = Normal;
            continue;
        }
        if (state == OutPath) {
            o->out_path = arg;
            state = Normal;
            continue;
        }
        if (arg.empty()) return fail("InvalidCommand", arg);
        if (arg[0] == '-') {
            if (arg.size() > 1 && arg[1] == '-') {
                const std::string name = arg.substr(2);
                if (name == "help") o->help = true;
                else if (name == "print") o->print = true;
                else if (name == "debug") o->debug = true;
                else if (name == "test") o->dry = true;
                else if (name == "strict") o->strict = true;
                else if (name == "output") state = OutPath;
                else return fail("InvalidOption", arg);
                if (o->help) return 0;
                continue;
            }
            for (size_t k = 1; k < arg.size(); ++k) {
                switch (arg[k]) {
                    case 'h': o->help = true; return 0;
                    case 'p': o->print = true; break;
                    case 'd': o->debug = true; break;
                    case 't': o->dry = true; break;
                    case 'o': state = OutPath; break;
                    default: return fail("InvalidOption", arg);
                }
            }
        } else if (arg[0] == 'c' || arg[0] == 'd') {  // main.zig:123-129: any word starting with c / d
            o->mode = arg[0] == 'c' ? Mode::Compress : Mode::Decompress;
            state = InPath;
        } else {
            return fail("InvalidCommand", arg);
        }
    }
    return 0;
}

std::string default_out_path(const Options &o) {
    if (o.mode == Mode::Compress) return o.in_path + ".et";
    const size_t slash = o.in_path.find_last_of('/');
    const std::string dir = slash == std::string::npos ? "" : o.in_path.substr(0, slash + 1);
    std::string base = slash == std::string::npos ? o.in_path : o.in_path.substr(slash + 1);
    if (base.size() >= 3 && base.compare(base.size() - 3, 3, ".et") == 0) base.resize(base.size() - 3);
    return dir + "decoded_" + base;
}

struct Pinned {
    uint8_t *p = nullptr;
    size_t n = 0;
    ~Pinned() { et_free_pinned(p); }
    bool alloc(size_t bytes) {
        n = bytes;
        void *q = nullptr;
        if (et_alloc_pinned(bytes ? bytes : 1, &q) != ET_OK) return false;
        p = static_cast<uint8_t *>(q);
        return true;
    }
};

}  // namespace

int main(int argc, char **argv) {
    Options o;
    if (int rc = parse(argc, argv, &o)) return rc;
    if (o.help || o.mode == Mode::None) {
        if (o.help || argc <= 1) std::fputs(kHelp, stdout);
        return 0;
    }
    if (o.in_path.empty()) return fail("NoInputFile", "");
    if (o.out_path.empty()) o.out_path = default_out_path(o);

    // main.zig:34-40: the whole file in one buffer
    FILE *f = std::fopen(o.in_path.c_str(), "rb");
    if (!f) return fail("FileNotFound", o.in_path + " (" + std::strerror(errno) + ")");
    std::fseek(f, 0, SEEK_END);
    const long size = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    Pinned in;
    if (size < 0 || !in.alloc((size_t)size)) {
        std::fclose(f);
        return fail("OutOfMemory", "input buffer (is a CUDA device present?)");
    }
    const size_t got = std::fread(in.p, 1, (size_t)size, f);
    std::fclose(f);
    if (got != (size_t)size) return fail("EndOfStream", o.in_path);

    et_ctx *ctx = nullptr;
    int rc = et_ctx_create(0, &ctx);
    if (rc != ET_OK) return fail(et_strerror(rc), "");
    et_ctx_set_output_fd(ctx, 1);
    const uint32_t flags = (o.dry ? 0u : ET_FLAG_WRITE_OUTPUT) | (o.print ? ET_FLAG_PRINT_OUTPUT : 0u) |
                           (o.debug ? ET_FLAG_DEBUG : 0u);
    Pinned out;
    size_t out_len = 0;
    if (o.mode == Mode::Compress) {
        // 7200 + n is the reference's scratch (encode.zig:253); a file whose codes average more
        // than 8 bits would overflow it there (NoSpaceLeft) and does here too
        if (!out.alloc(et_encode_bound(in.n))) rc = ET_ERR_OUT_OF_MEMORY;
        else rc = et_encode(ctx, in.p, in.n, out.p, out.n, &out_len, flags);
    } else {
        if (in.n < 9) rc = ET_ERR_CORRUPT;
        else if (o.strict && !(in.p[0] == 0xe7 && in.p[1] == 0xc0 && in.p[2] == 0xde && in.p[3] == 0x01)) rc = ET_ERR_CORRUPT;
        else {
            // body length: BE u32 after the entry count (decode.zig:36-42); main.zig:204 passes file[4..]
            const size_t n_text = ((size_t)in.p[5] << 24) | ((size_t)in.p[6] << 16) | ((size_t)in.p[7] << 8) | in.p[8];
            if (!out.alloc(n_text)) rc = ET_ERR_OUT_OF_MEMORY;
            else rc = et_decode(ctx, in.p + 4, in.n - 4, out.p, out.n, &out_len, flags);
        }
    }
    if (rc != ET_OK) {
        const std::string detail = et_last_error(ctx);
        et_ctx_destroy(ctx);
        return fail(et_strerror(rc), detail);
    }
    et_ctx_destroy(ctx);
    if (!o.dry) {  // main.zig:191-197
        FILE *g = std::fopen(o.out_path.c_str(), "wb");
        if (!g) return fail("AccessDenied", o.out_path + " (" + std::strerror(errno) + ")");
        const size_t w = std::fwrite(out.p, 1, out_len, g);
        std::fclose(g);
        if (w != out_len) return fail("NoSpaceLeft", o.out_path);
    }
    return 0;
}
