// Dependency-free reader of PCD v0.7 point-cloud files — what the node feeds its map from
// (pcl::io::loadPCDFile<PointType>, pcm_matching/src/pcm_matching.cpp:69-79; PointType = pcl::PointXYZINormal, of which the
// map only ever uses x, y, z: Pcl2PointStruct, pcm_matching.hpp:205-220).  PCL is a third-party dependency that is not in the
// reference tree; this restates the published file format:
//   header lines  VERSION / FIELDS / SIZE / TYPE / COUNT / WIDTH / HEIGHT / VIEWPOINT / POINTS / DATA
//   DATA ascii              one point per line, fields separated by blanks
//   DATA binary             POINTS packed records in FIELDS order, little endian
//   DATA binary_compressed  uint32 compressed size, uint32 uncompressed size, LZF stream; uncompressed layout is
//                           field-major (all x, then all y, ...)
// Only x, y, z are extracted (as float32: TYPE F SIZE 4, or converted from F 8 / I / U).  Points with a non-finite
// coordinate (PCL writes NaN for invalid returns) are dropped — the reference would feed them to static_cast<int>.
#include "pcd_reader.hpp"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <sstream>

namespace elm {

namespace {

struct FileCloser { void operator()(std::FILE* f) const { if (f) std::fclose(f); } };

struct Field { std::string name; int size = 4; char type = 'F'; int count = 1; size_t offset = 0; };

struct Header {
    std::vector<Field> fields;
    size_t points = 0, width = 0, height = 1, record = 0;
    std::string data;
    int ix = -1, iy = -1, iz = -1;
};

// one header line without the trailing newline; false at EOF
bool read_line(std::FILE* f, std::string& line) {
    line.clear();
    int c;
    while ((c = std::fgetc(f)) != EOF) {
        if (c == '\n') return true;
        if (c != '\r') line.push_back(static_cast<char>(c));
    }
    return !line.empty();
}

std::string parse_header(std::FILE* f, Header& h) {
    std::string line;
    bool have_points = false;
    while (read_line(f, line)) {
        if (line.empty() || line[0] == '#') continue;
        std::istringstream ss(line);
        std::string key;
        ss >> key;
        if (key == "VERSION" || key == "VIEWPOINT") continue;
        if (key == "FIELDS") { std::string n; while (ss >> n) { Field fd; fd.name = n; h.fields.push_back(fd); } }
        else if (key == "SIZE") { for (Field& fd : h.fields) if (!(ss >> fd.size)) return "PCD header: SIZE shorter than FIELDS"; }
        else if (key == "TYPE") { for (Field& fd : h.fields) if (!(ss >> fd.type)) return "PCD header: TYPE shorter than FIELDS"; }
        else if (key == "COUNT") { for (Field& fd : h.fields) if (!(ss >> fd.count)) return "PCD header: COUNT shorter than FIELDS"; }
        else if (key == "WIDTH") ss >> h.width;
        else if (key == "HEIGHT") ss >> h.height;
        else if (key == "POINTS") { ss >> h.points; have_points = true; }
        else if (key == "DATA") { ss >> h.data; break; }
        else return "PCD header: unknown entry '" + key + "'";
    }
    if (h.data.empty()) return "PCD header: no DATA line";
    if (h.fields.empty()) return "PCD header: no FIELDS line";
    if (!have_points) h.points = h.width * h.height;
    size_t off = 0;
    for (size_t i = 0; i < h.fields.size(); ++i) {
        Field& fd = h.fields[i];
        if (fd.size != 1 && fd.size != 2 && fd.size != 4 && fd.size != 8) return "PCD header: unsupported SIZE";
        if (fd.type != 'F' && fd.type != 'I' && fd.type != 'U') return "PCD header: unsupported TYPE";
        if (fd.count < 1 || fd.count > 65536) return "PCD header: COUNT out of range";
        fd.offset = off;
        off += static_cast<size_t>(fd.size) * static_cast<size_t>(fd.count);
        if (off > (1u << 16)) return "PCD header: record larger than 64 KiB";  // (keeps POINTS x record far from overflowing size_t)
        if (fd.name == "x") h.ix = static_cast<int>(i);
        if (fd.name == "y") h.iy = static_cast<int>(i);
        if (fd.name == "z") h.iz = static_cast<int>(i);
    }
    h.record = off;
    if (h.ix < 0 || h.iy < 0 || h.iz < 0) return "PCD file has no x / y / z fields";
    if (h.points > (1ull << 32)) return "PCD file: more than 2^32 points";
    return "";
}

float decode(const unsigned char* p, const Field& fd) {
    switch (fd.type) {
        case 'F':
            if (fd.size == 4) { float v; std::memcpy(&v, p, 4); return v; }
            if (fd.size == 8) { double v; std::memcpy(&v, p, 8); return static_cast<float>(v); }
            return NAN;
        case 'I':
            if (fd.size == 1) { int8_t v; std::memcpy(&v, p, 1); return v; }
            if (fd.size == 2) { int16_t v; std::memcpy(&v, p, 2); return v; }
            if (fd.size == 4) { int32_t v; std::memcpy(&v, p, 4); return static_cast<float>(v); }
            { int64_t v; std::memcpy(&v, p, 8); return static_cast<float>(v); }
        default:
            if (fd.size == 1) { uint8_t v; std::memcpy(&v, p, 1); return v; }
            if (fd.size == 2) { uint16_t v; std::memcpy(&v, p, 2); return v; }
            if (fd.size == 4) { uint32_t v; std::memcpy(&v, p, 4); return static_cast<float>(v); }
            { uint64_t v; std::memcpy(&v, p, 8); return static_cast<float>(v); }
    }
}

void push_if_finite(std::vector<float>& xyz, float x, float y, float z, size_t& dropped) {
    if (std::isfinite(x) && std::isfinite(y) && std::isfinite(z)) { xyz.push_back(x); xyz.push_back(y); xyz.push_back(z); }
    else ++dropped;
}

// LZF decompression (Marc Lehmann's liblzf format, the codec of PCD's binary_compressed): a control byte < 32 starts a
// literal run of ctrl + 1 bytes; otherwise it is a back reference of length (ctrl >> 5) + 2 (length field 7: one more length
// byte follows) at distance ((ctrl & 31) << 8 | next byte) + 1.
bool lzf_decompress(const unsigned char* in, size_t in_len, unsigned char* out, size_t out_len) {
    size_t ip = 0, op = 0;
    while (ip < in_len) {
        unsigned ctrl = in[ip++];
        if (ctrl < 32) {
            ++ctrl;
            if (ip + ctrl > in_len || op + ctrl > out_len) return false;
            std::memcpy(out + op, in + ip, ctrl);
            ip += ctrl; op += ctrl;
        } else {
            size_t len = ctrl >> 5;
            if (len == 7) { if (ip >= in_len) return false; len += in[ip++]; }
            if (ip >= in_len) return false;
            const size_t dist = ((static_cast<size_t>(ctrl) & 31u) << 8 | in[ip++]) + 1;
            len += 2;
            if (dist > op || op + len > out_len) return false;
            for (size_t k = 0; k < len; ++k, ++op) out[op] = out[op - dist];  // may overlap: byte by byte
        }
    }
    return op == out_len;
}

}  // namespace

std::string read_pcd_xyz(const std::string& path, std::vector<float>& xyz, size_t* dropped_out) {
    xyz.clear();
    size_t dropped = 0;
    std::unique_ptr<std::FILE, FileCloser> f(std::fopen(path.c_str(), "rb"));
    if (!f) return "cannot open " + path;
    Header h;
    const std::string e = parse_header(f.get(), h);
    if (!e.empty()) return path + ": " + e;
    const Field &fx = h.fields[h.ix], &fy = h.fields[h.iy], &fz = h.fields[h.iz];
    // POINTS comes from an untrusted header: never reserve more than the file can hold (a point needs at least 6 bytes in
    // ascii — three one-digit values, separators, newline — and `record` bytes in the binary encodings, where LZF expands
    // at most ~264 x), so a forged count cannot exhaust the host's memory before the data turns out to be missing
    {
        const long pos = std::ftell(f.get());
        size_t remaining = 0;
        if (pos >= 0 && std::fseek(f.get(), 0, SEEK_END) == 0) {
            const long end = std::ftell(f.get());
            if (end >= pos) remaining = static_cast<size_t>(end - pos);
            std::fseek(f.get(), pos, SEEK_SET);
        }
        const size_t per_point = h.data == "ascii" ? 6 : h.record;
        const size_t expansion = h.data == "binary_compressed" ? 264 : 1;
        const size_t fit = per_point ? (remaining / per_point + 1) * expansion : 0;
        if (h.points > fit) return path + ": POINTS (" + std::to_string(h.points) + ") exceeds what the file can hold";
    }
    xyz.reserve(3 * h.points);
    if (h.data == "ascii") {
        std::string line;
        std::vector<double> vals;
        size_t col_x = 0, col_y = 0, col_z = 0, cols = 0;
        for (size_t i = 0; i < h.fields.size(); ++i) {
            if (static_cast<int>(i) == h.ix) col_x = cols;
            if (static_cast<int>(i) == h.iy) col_y = cols;
            if (static_cast<int>(i) == h.iz) col_z = cols;
            cols += static_cast<size_t>(h.fields[i].count);
        }
        for (size_t p = 0; p < h.points; ++p) {
            if (!read_line(f.get(), line)) return path + ": ascii data ends after " + std::to_string(p) + " of " + std::to_string(h.points) + " points";
            vals.clear();
            const char* s = line.c_str();
            char* end = nullptr;
            for (;;) {
                const double v = std::strtod(s, &end);  // accepts "nan"
                if (end == s) break;
                vals.push_back(v);
                s = end;
            }
            if (vals.size() < cols) return path + ": ascii record " + std::to_string(p) + " has too few values";
            push_if_finite(xyz, static_cast<float>(vals[col_x]), static_cast<float>(vals[col_y]), static_cast<float>(vals[col_z]), dropped);
        }
    } else if (h.data == "binary") {
        std::vector<unsigned char> buf(std::min<size_t>(h.points, 1u << 16) * h.record);
        for (size_t done = 0; done < h.points;) {
            const size_t n = std::min<size_t>(h.points - done, 1u << 16);
            if (std::fread(buf.data(), h.record, n, f.get()) != n) return path + ": binary data truncated";
            for (size_t i = 0; i < n; ++i) {
                const unsigned char* r = buf.data() + i * h.record;
                push_if_finite(xyz, decode(r + fx.offset, fx), decode(r + fy.offset, fy), decode(r + fz.offset, fz), dropped);
            }
            done += n;
        }
    } else if (h.data == "binary_compressed") {
        uint32_t csize = 0, usize = 0;
        if (std::fread(&csize, 4, 1, f.get()) != 1 || std::fread(&usize, 4, 1, f.get()) != 1) return path + ": compressed sizes missing";
        if (static_cast<size_t>(usize) != h.points * h.record) return path + ": uncompressed size does not match POINTS x record size";
        // (record <= 64 KiB and POINTS <= 2^32: the product above cannot wrap, and every field's plane
        //  [points * offset, points * (offset + size * count)) lies inside the usize bytes)
        std::vector<unsigned char> cbuf(csize), ubuf(usize);
        if (csize && std::fread(cbuf.data(), 1, csize, f.get()) != csize) return path + ": compressed data truncated";
        if (!lzf_decompress(cbuf.data(), csize, ubuf.data(), usize)) return path + ": corrupt LZF stream";
        // field-major: field k occupies [points * offset_k, points * (offset_k + size_k * count_k))
        const unsigned char* bx = ubuf.data() + h.points * fx.offset;
        const unsigned char* by = ubuf.data() + h.points * fy.offset;
        const unsigned char* bz = ubuf.data() + h.points * fz.offset;
        for (size_t i = 0; i < h.points; ++i)
            push_if_finite(xyz, decode(bx + i * fx.size * fx.count, fx), decode(by + i * fy.size * fy.count, fy), decode(bz + i * fz.size * fz.count, fz),
                           dropped);
    } else {
        return path + ": unsupported DATA '" + h.data + "'";
    }
    if (dropped_out) *dropped_out = dropped;
    return "";
}

}  // namespace elm
