// ParaView output — replaces new_pvd_file / save_frame! / save_pvd_file
// (src/IO.jl:20-75), which delegate to WriteVTK.jl.  File layout follows what
// WriteVTK emits for `vtk_grid(path, points, [MeshCell(PolyData.Verts(), [i])...])`,
// confirmed on the reference's fixture sph_jl/examples/init/cylinder.vtp:
//   <VTKFile type="PolyData" version="1.0" byte_order="LittleEndian"
//            header_type="UInt64" compressor="vtkZLibDataCompressor">
//   Points (Float64 x3), Verts connectivity/offsets (Int64), one Float64
//   PointData array per exported field, all format="appended", raw encoding,
//   each array = [nblocks, blocksize, lastblocksize, csize...] + zlib blocks.
#include <errno.h>
#include <stdio.h>
#include <string.h>
#include <sys/stat.h>
#include <zlib.h>

#include <stdlib.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "sphmw_internal.h"

namespace {

const size_t BLOCK = 1u << 15;  // vtkZLibDataCompressor default block size (32 KiB)

// one appended array: UInt64 header + compressed blocks.  The blocks are independent zlib
// streams, so they are compressed by a few host threads (a 64 M-particle frame is ~6 GB of
// doubles; one thread at zlib level 1 would take a minute).  SPHMW_IO_THREADS overrides.
struct AppendedArray {
    std::vector<uint64_t> header;      // [nblocks, blocksize, lastblocksize, csize...]
    std::vector<std::string> parts;    // compressed blocks, one string per thread, in order
    size_t bytes() const {
        size_t b = header.size() * sizeof(uint64_t);
        for (const std::string &p : parts) b += p.size();
        return b;
    }
};

unsigned io_threads(size_t nblocks) {
    unsigned t = std::thread::hardware_concurrency();
    if (const char *e = getenv("SPHMW_IO_THREADS")) t = (unsigned)atoi(e);
    if (t < 1) t = 1;
    if (t > 64) t = 64;
    const size_t by_work = nblocks / 64 + 1;  // at least ~2 MB of input per thread
    return (unsigned)std::min<size_t>(t, by_work);
}

AppendedArray compress_array(const void *data, size_t nbytes) {
    const unsigned char *src = (const unsigned char *)data;
    const size_t nblocks = nbytes == 0 ? 0 : (nbytes + BLOCK - 1) / BLOCK;
    size_t last = nbytes == 0 ? 0 : nbytes - (nblocks - 1) * BLOCK;
    if (last == BLOCK) last = 0;  // VTK convention: 0 means "last block is full"
    AppendedArray a;
    a.header.resize(3 + nblocks);
    a.header[0] = nblocks;
    a.header[1] = BLOCK;
    a.header[2] = last;
    const unsigned nt = io_threads(nblocks);
    a.parts.resize(nt);
    auto work = [&](unsigned t) {
        const size_t b0 = nblocks * t / nt, b1 = nblocks * (t + 1) / nt;
        std::vector<unsigned char> buf(compressBound(BLOCK));
        std::string &out = a.parts[t];
        out.reserve((b1 - b0) * BLOCK / 2);
        for (size_t b = b0; b < b1; ++b) {
            const size_t len = std::min(BLOCK, nbytes - b * BLOCK);
            uLongf clen = buf.size();
            compress2(buf.data(), &clen, src + b * BLOCK, len, 1);
            a.header[3 + b] = clen;
            out.append((const char *)buf.data(), clen);
        }
    };
    if (nt <= 1) {
        work(0);
    } else {
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < nt; ++t) pool.emplace_back(work, t);
        for (std::thread &th : pool) th.join();
    }
    return a;
}

bool write_array(FILE *fp, const AppendedArray &a) {
    if (fwrite(a.header.data(), sizeof(uint64_t), a.header.size(), fp) != a.header.size()) return false;
    for (const std::string &p : a.parts)
        if (!p.empty() && fwrite(p.data(), 1, p.size(), fp) != p.size()) return false;
    return true;
}

int mkpath(const std::string &path) {
    std::string cur;
    for (size_t i = 0; i <= path.size(); ++i) {
        if (i == path.size() || path[i] == '/') {
            if (!cur.empty() && mkdir(cur.c_str(), 0777) != 0 && errno != EEXIST) return -1;
        }
        if (i < path.size()) cur.push_back(path[i]);
    }
    return 0;
}

}  // namespace

// points3n: xyz interleaved per point (the 3xN column-major matrix of IO.jl:39-42)
// data[f]: ncomps[f] x N column-major, i.e. interleaved per point (IO.jl:62-68)
int sphmw_write_vtp(const char *path, int64_t n, const double *points3n, int nfields,
                    const char *const *names, const int *ncomps, const double *const *data) {
    std::vector<AppendedArray> arrays;
    arrays.push_back(compress_array(points3n, sizeof(double) * 3 * (size_t)n));
    {
        std::vector<int64_t> conn((size_t)n), offs((size_t)n);
        for (int64_t i = 0; i < n; ++i) {
            conn[i] = i;
            offs[i] = i + 1;
        }
        arrays.push_back(compress_array(conn.data(), sizeof(int64_t) * (size_t)n));
        arrays.push_back(compress_array(offs.data(), sizeof(int64_t) * (size_t)n));
    }
    for (int f = 0; f < nfields; ++f)
        arrays.push_back(compress_array(data[f], sizeof(double) * (size_t)ncomps[f] * (size_t)n));
    std::vector<size_t> offsets;
    size_t running = 0;
    for (const AppendedArray &a : arrays) {
        offsets.push_back(running);
        running += a.bytes();
    }

    FILE *fp = fopen(path, "wb");
    if (!fp) {
        sphmw_set_error("cannot open %s: %s", path, strerror(errno));
        return SPHMW_E_IO;
    }
    fprintf(fp, "<?xml version=\"1.0\" encoding=\"utf-8\"?>\n");
    fprintf(fp,
            "<VTKFile type=\"PolyData\" version=\"1.0\" byte_order=\"LittleEndian\" "
            "header_type=\"UInt64\" compressor=\"vtkZLibDataCompressor\">\n");
    fprintf(fp, "  <PolyData>\n");
    fprintf(fp, "    <Piece NumberOfPoints=\"%lld\" NumberOfVerts=\"%lld\">\n", (long long)n,
            (long long)n);
    fprintf(fp, "      <Points>\n");
    fprintf(fp,
            "        <DataArray type=\"Float64\" Name=\"Points\" NumberOfComponents=\"3\" "
            "format=\"appended\" offset=\"%zu\"/>\n",
            offsets[0]);
    fprintf(fp, "      </Points>\n");
    fprintf(fp, "      <Verts>\n");
    fprintf(fp,
            "        <DataArray type=\"Int64\" Name=\"connectivity\" NumberOfComponents=\"1\" "
            "format=\"appended\" offset=\"%zu\"/>\n",
            offsets[1]);
    fprintf(fp,
            "        <DataArray type=\"Int64\" Name=\"offsets\" NumberOfComponents=\"1\" "
            "format=\"appended\" offset=\"%zu\"/>\n",
            offsets[2]);
    fprintf(fp, "      </Verts>\n");
    fprintf(fp, "      <PointData>\n");
    for (int f = 0; f < nfields; ++f)
        fprintf(fp,
                "        <DataArray type=\"Float64\" Name=\"%s\" NumberOfComponents=\"%d\" "
                "format=\"appended\" offset=\"%zu\"/>\n",
                names[f], ncomps[f], offsets[3 + f]);
    fprintf(fp, "      </PointData>\n");
    fprintf(fp, "    </Piece>\n");
    fprintf(fp, "  </PolyData>\n");
    fprintf(fp, "  <AppendedData encoding=\"raw\">\n_");
    bool ok = true;
    for (const AppendedArray &a : arrays) ok = ok && write_array(fp, a);
    fprintf(fp, "\n  </AppendedData>\n</VTKFile>\n");
    if (fclose(fp) != 0 || !ok) {
        sphmw_set_error("write to %s failed", path);
        return SPHMW_E_IO;
    }
    return SPHMW_OK;
}

// paraview_collection: one <DataSet timestep=frame index> per frame (IO.jl:73)
int sphmw_write_pvd(const char *path, const std::vector<std::string> &files) {
    FILE *fp = fopen(path, "wb");
    if (!fp) {
        sphmw_set_error("cannot open %s: %s", path, strerror(errno));
        return SPHMW_E_IO;
    }
    fprintf(fp, "<?xml version=\"1.0\" encoding=\"utf-8\"?>\n");
    fprintf(fp, "<VTKFile type=\"Collection\" version=\"1.0\" byte_order=\"LittleEndian\">\n");
    fprintf(fp, "  <Collection>\n");
    for (size_t i = 0; i < files.size(); ++i)
        fprintf(fp, "    <DataSet timestep=\"%zu.0\" part=\"0\" file=\"%s\"/>\n", i, files[i].c_str());
    fprintf(fp, "  </Collection>\n</VTKFile>\n");
    fclose(fp);
    return SPHMW_OK;
}

// ≙ new_pvd_file(path) — IO.jl:20-26
extern "C" int sphmw_pvd_open(sphmw_ctx *c, const char *dir) {
    if (!c || !dir) { sphmw_set_error("null argument"); return SPHMW_E_INVALID; }
    if (mkpath(dir) != 0) {
        sphmw_set_error("cannot create %s: %s", dir, strerror(errno));
        return SPHMW_E_IO;
    }
    c->pvd_dir = dir;
    c->pvd_frame = 0;
    c->pvd_entries.clear();
    c->pvd_open = true;
    return SPHMW_OK;
}

// ≙ save_frame!(data, sys, vars...) — IO.jl:53-75 (+ capture_frame :37-46).  The frame is
// captured on the device at this point of the stream (frame_async.cu: index order, components
// interleaved), copied to pinned host memory on a side stream and written by a worker thread; the
// call returns as soon as the capture is queued, so the time loop goes on while the file is made.
extern "C" int sphmw_pvd_save_frame(sphmw_ctx *c, const char *const *fields, int32_t nfields) {
    if (!c || (nfields > 0 && !fields)) { sphmw_set_error("null argument"); return SPHMW_E_INVALID; }
    if (!c->pvd_open) { sphmw_set_error("save_frame: no pvd file is open"); return SPHMW_E_STATE; }
    if (cudaSetDevice(c->device) != cudaSuccess) { sphmw_set_error("cudaSetDevice failed"); return SPHMW_E_CUDA; }
    std::string fname = "frame" + std::to_string(c->pvd_frame) + ".vtp";
    std::string path = c->pvd_dir + "/" + fname;
    TRY(sphmw_pvd_save_frame_async(c, fields, nfields, path));
    c->pvd_entries.push_back(fname);
    c->pvd_frame += 1;
    return SPHMW_OK;
}

// ≙ save_pvd_file(data) — IO.jl:33-35
extern "C" int sphmw_pvd_close(sphmw_ctx *c) {
    if (!c) return SPHMW_E_INVALID;
    if (!c->pvd_open) { sphmw_set_error("save_pvd_file: no pvd file is open"); return SPHMW_E_STATE; }
    TRY(sphmw_frame_async_drain(c));  // every frame is on disk before the collection names it
    std::string path = c->pvd_dir + "/result.pvd";
    TRY(sphmw_write_pvd(path.c_str(), c->pvd_entries));
    c->pvd_open = false;
    return SPHMW_OK;
}

// ===========================================================================
// .vtp reader — the file-format half of import_particles! (src/IO.jl:83-122, which
// delegates to ReadVTK.jl).  Reads what WriteVTK writes (and what this file writes):
// PolyData, appended raw data, optional vtkZLibDataCompressor, UInt64/UInt32 headers.
// Host-only: no CUDA context is needed.
// ===========================================================================
struct VtpArray {
    std::string name, type;
    int ncomp = 1;
    size_t offset = 0;
};
struct sphmw_vtp {
    int64_t n_points = 0;
    bool compressed = false;
    int header_bytes = 8;
    std::vector<VtpArray> arrays;  // [0] is "Points"
    std::string blob;              // appended data after the '_'
};

namespace {
std::string attr(const std::string &tag, const char *name) {
    std::string key = std::string(name) + "=\"";
    size_t a = tag.find(key);
    if (a == std::string::npos) return "";
    a += key.size();
    size_t b = tag.find('"', a);
    return tag.substr(a, b - a);
}
uint64_t rd_uint(const std::string &blob, size_t pos, int bytes) {
    uint64_t v = 0;
    memcpy(&v, blob.data() + pos, bytes);
    return v;
}
size_t type_size(const std::string &t) {
    if (t == "Float64" || t == "Int64" || t == "UInt64") return 8;
    if (t == "Float32" || t == "Int32" || t == "UInt32") return 4;
    if (t == "Int16" || t == "UInt16") return 2;
    if (t == "Int8" || t == "UInt8") return 1;
    return 0;
}
double to_double(const std::string &t, const unsigned char *p) {
    if (t == "Float64") { double v; memcpy(&v, p, 8); return v; }
    if (t == "Float32") { float v; memcpy(&v, p, 4); return v; }
    if (t == "Int64") { int64_t v; memcpy(&v, p, 8); return (double)v; }
    if (t == "UInt64") { uint64_t v; memcpy(&v, p, 8); return (double)v; }
    if (t == "Int32") { int32_t v; memcpy(&v, p, 4); return v; }
    if (t == "UInt32") { uint32_t v; memcpy(&v, p, 4); return v; }
    if (t == "Int16") { int16_t v; memcpy(&v, p, 2); return v; }
    if (t == "UInt16") { uint16_t v; memcpy(&v, p, 2); return v; }
    if (t == "Int8") { int8_t v; memcpy(&v, p, 1); return v; }
    uint8_t v; memcpy(&v, p, 1); return v;
}
}  // namespace

extern "C" int sphmw_vtp_open(const char *path, sphmw_vtp **out) {
    if (!path || !out) { sphmw_set_error("null argument"); return SPHMW_E_INVALID; }
    *out = nullptr;
    FILE *fp = fopen(path, "rb");
    if (!fp) { sphmw_set_error("cannot open %s: %s", path, strerror(errno)); return SPHMW_E_IO; }
    std::string raw;
    char buf[1 << 16];
    size_t k;
    while ((k = fread(buf, 1, sizeof(buf), fp)) > 0) raw.append(buf, k);
    fclose(fp);
    size_t app = raw.find("<AppendedData");
    if (app == std::string::npos) { sphmw_set_error("%s: no <AppendedData> section (only appended raw .vtp files are supported)", path); return SPHMW_E_IO; }
    std::string head = raw.substr(0, app);
    size_t us = raw.find('_', raw.find('>', app));
    if (us == std::string::npos) { sphmw_set_error("%s: malformed appended data", path); return SPHMW_E_IO; }
    sphmw_vtp *v = new sphmw_vtp();
    {
        size_t a = head.find("<VTKFile");
        std::string tag = head.substr(a, head.find('>', a) - a);
        if (attr(tag, "type") != "PolyData") { delete v; sphmw_set_error("%s: not a PolyData file", path); return SPHMW_E_IO; }
        v->compressed = !attr(tag, "compressor").empty();
        v->header_bytes = attr(tag, "header_type") == "UInt64" ? 8 : 4;
        std::string enc = raw.substr(app, raw.find('>', app) - app);
        if (attr(enc, "encoding") != "raw") { delete v; sphmw_set_error("%s: only raw appended data is supported", path); return SPHMW_E_IO; }
        a = head.find("<Piece");
        tag = head.substr(a, head.find('>', a) - a);
        v->n_points = atoll(attr(tag, "NumberOfPoints").c_str());
    }
    size_t pd0 = head.find("<PointData"), pd1 = head.find("</PointData>");
    size_t pt0 = head.find("<Points"), pt1 = head.find("</Points>");
    size_t pos = 0;
    VtpArray points;
    std::vector<VtpArray> fields;
    while ((pos = head.find("<DataArray", pos)) != std::string::npos) {
        size_t end = head.find('>', pos);
        std::string tag = head.substr(pos, end - pos);
        VtpArray a;
        a.name = attr(tag, "Name");
        a.type = attr(tag, "type");
        std::string nc = attr(tag, "NumberOfComponents");
        a.ncomp = nc.empty() ? 1 : atoi(nc.c_str());
        a.offset = (size_t)atoll(attr(tag, "offset").c_str());
        if (attr(tag, "format") != "appended" || type_size(a.type) == 0) {
            delete v;
            sphmw_set_error("%s: DataArray %s is not appended data of a known type", path, a.name.c_str());
            return SPHMW_E_IO;
        }
        if (pt0 != std::string::npos && pos > pt0 && pos < pt1) points = a;
        else if (pd0 != std::string::npos && pos > pd0 && pos < pd1) fields.push_back(a);
        pos = end;
    }
    if (points.type.empty()) { delete v; sphmw_set_error("%s: no Points array", path); return SPHMW_E_IO; }
    points.name = "Points";
    v->arrays.push_back(points);
    for (auto &f : fields) v->arrays.push_back(f);
    v->blob = raw.substr(us + 1);
    *out = v;
    return SPHMW_OK;
}

extern "C" int sphmw_vtp_close(sphmw_vtp *v) {
    delete v;
    return SPHMW_OK;
}

extern "C" int sphmw_vtp_info(sphmw_vtp *v, int64_t *n_points, int32_t *n_arrays) {
    if (!v) return SPHMW_E_INVALID;
    if (n_points) *n_points = v->n_points;
    if (n_arrays) *n_arrays = (int32_t)v->arrays.size();
    return SPHMW_OK;
}

// i = 0 is "Points" (3 components); the PointData arrays follow in file order
extern "C" int sphmw_vtp_array(sphmw_vtp *v, int32_t i, char *name, int64_t cap, int32_t *ncomp) {
    if (!v || i < 0 || i >= (int32_t)v->arrays.size()) { sphmw_set_error("array index out of range"); return SPHMW_E_INVALID; }
    if (name && cap > 0) {
        size_t k = std::min<size_t>(v->arrays[i].name.size(), (size_t)cap - 1);
        memcpy(name, v->arrays[i].name.data(), k);
        name[k] = 0;
    }
    if (ncomp) *ncomp = v->arrays[i].ncomp;
    return SPHMW_OK;
}

// values as doubles, interleaved per point (ncomp x N column-major, as the file stores them)
extern "C" int sphmw_vtp_read(sphmw_vtp *v, const char *name, double *out, int64_t n_values) {
    if (!v || !name || !out) { sphmw_set_error("null argument"); return SPHMW_E_INVALID; }
    const VtpArray *a = nullptr;
    for (auto &x : v->arrays)
        if (x.name == name) a = &x;
    if (!a) { sphmw_set_error("Variable %s does not exist!", name); return SPHMW_E_UNKNOWN_FIELD; }
    const size_t ts = type_size(a->type);
    const size_t want = (size_t)v->n_points * a->ncomp;
    if ((size_t)n_values != want) { sphmw_set_error("vtp_read(%s): expected %zu values", name, want); return SPHMW_E_INVALID; }
    std::vector<unsigned char> bytes;
    const int hb = v->header_bytes;
    size_t pos = a->offset;
    if (pos + 3 * (size_t)hb > v->blob.size() && v->compressed) { sphmw_set_error("truncated file"); return SPHMW_E_IO; }
    if (v->compressed) {
        uint64_t nb = rd_uint(v->blob, pos, hb), bs = rd_uint(v->blob, pos + hb, hb), last = rd_uint(v->blob, pos + 2 * hb, hb);
        size_t cpos = pos + (3 + nb) * hb;
        for (uint64_t b = 0; b < nb; ++b) {
            uint64_t cs = rd_uint(v->blob, pos + (3 + b) * hb, hb);
            uLongf len = (uLongf)((b + 1 == nb && last) ? last : bs);
            size_t old = bytes.size();
            bytes.resize(old + len);
            if (cpos + cs > v->blob.size() ||
                uncompress(bytes.data() + old, &len, (const Bytef *)v->blob.data() + cpos, (uLong)cs) != Z_OK) {
                sphmw_set_error("vtp_read(%s): corrupt compressed block", name);
                return SPHMW_E_IO;
            }
            bytes.resize(old + len);
            cpos += cs;
        }
    } else {
        uint64_t nbytes = rd_uint(v->blob, pos, hb);
        if (pos + hb + nbytes > v->blob.size()) { sphmw_set_error("truncated file"); return SPHMW_E_IO; }
        bytes.assign((const unsigned char *)v->blob.data() + pos + hb, (const unsigned char *)v->blob.data() + pos + hb + nbytes);
    }
    if (bytes.size() != want * ts) {
        sphmw_set_error("vtp_read(%s): %zu bytes, expected %zu", name, bytes.size(), want * ts);
        return SPHMW_E_IO;
    }
    for (size_t i = 0; i < want; ++i) out[i] = to_double(a->type, bytes.data() + i * ts);
    return SPHMW_OK;
}

// stand-alone writer (no context): the same file save_frame! produces
extern "C" int sphmw_vtp_write(const char *path, int64_t n, const double *points3n, int32_t nfields,
                               const char *const *names, const int32_t *ncomps,
                               const double *const *data) {
    if (!path || (n > 0 && !points3n)) { sphmw_set_error("null argument"); return SPHMW_E_INVALID; }
    std::vector<int> nc(ncomps, ncomps + nfields);
    return sphmw_write_vtp(path, n, points3n, nfields, names, nc.data(), data);
}
