// Dataset ingest at GPU speed: the reference re-parses the topology file on every
// reset (environment16.py:209) and one `.demands` file per episode with a Python line loop
// (defo_process_results.py:216-225); once episodes run in device-resident batches of thousands that loop
// is the bottleneck.  Same accept rules as the reference parser: skip two header lines, split every
// other line on single spaces, tm[int(c[1])][int(c[2])] = floor(float(c[3])).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "enero_internal.h"

namespace {

// returns "" on success, else an error message
std::string parse_demands(const char* path, int N, double* tm) {
    FILE* fp = std::fopen(path, "rb");
    if (!fp) return std::string("cannot open ") + path;
    std::fseek(fp, 0, SEEK_END);
    const long sz = std::ftell(fp);
    std::fseek(fp, 0, SEEK_SET);
    std::string buf((size_t)std::max<long>(sz, 0), '\0');
    if (sz > 0 && std::fread(&buf[0], 1, (size_t)sz, fp) != (size_t)sz) { std::fclose(fp); return std::string("short read of ") + path; }
    std::fclose(fp);
    std::memset(tm, 0, sizeof(double) * (size_t)N * N);
    const char* p = buf.c_str();
    const char* end = p + buf.size();
    int line_no = 0;
    while (p < end) {
        const char* eol = static_cast<const char*>(std::memchr(p, '\n', (size_t)(end - p)));
        if (!eol) eol = end;
        if (line_no >= 2) {
            // tokens separated by single spaces (Python's split(" "))
            const char* tok[4]; int ntok = 0;
            const char* q = p;
            tok[ntok++] = q;
            while (q < eol && ntok < 4) { if (*q == ' ') tok[ntok++] = q + 1; ++q; }
            if (ntok >= 4) {
                char* e1; char* e2; char* e3;
                const long s = std::strtol(tok[1], &e1, 10);
                const long d = std::strtol(tok[2], &e2, 10);
                const double v = std::strtod(tok[3], &e3);
                if (e1 == tok[1] || e2 == tok[2] || e3 == tok[3])
                    return std::string(path) + ": malformed demand on line " + std::to_string(line_no + 1);
                if (s < 0 || s >= N || d < 0 || d >= N)
                    return std::string(path) + ": node id out of range on line " + std::to_string(line_no + 1);
                tm[(size_t)s * N + d] = std::floor(v);
            }
        }
        ++line_no;
        p = eol + 1;
    }
    return "";
}

}  // namespace

extern "C" int enero_parse_demands(const char* path, int n_nodes, double* tm) {
    if (!path || !tm || n_nodes < 1) { enero_set_error("enero_parse_demands: bad argument"); return ENERO_ERR_INVALID; }
    const std::string err = parse_demands(path, n_nodes, tm);
    if (!err.empty()) { enero_set_error(err); return ENERO_ERR_INVALID; }
    return ENERO_OK;
}

extern "C" int enero_parse_demands_batch(const char* const* paths, int n_files, int n_nodes, double* tms, int n_threads) {
    if (!paths || !tms || n_files < 0 || n_nodes < 1) { enero_set_error("enero_parse_demands_batch: bad argument"); return ENERO_ERR_INVALID; }
    if (n_threads < 1) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    n_threads = std::min(n_threads, std::max(1, n_files));
    std::vector<std::string> errs((size_t)n_threads);
    auto work = [&](int t) {
        for (int i = t; i < n_files; i += n_threads) {
            const std::string e = parse_demands(paths[i], n_nodes, tms + (size_t)i * n_nodes * n_nodes);
            if (!e.empty() && errs[t].empty()) errs[t] = e;
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    for (const auto& e : errs)
        if (!e.empty()) { enero_set_error(e); return ENERO_ERR_INVALID; }
    return ENERO_OK;
}
