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

#include <string>
#include <vector>

#include "sphmw_internal.h"

namespace {

const size_t BLOCK = 1u << 15;  // vtkZLibDataCompressor default block size (32 KiB)

// one appended array: UInt64 header + compressed blocks
void append_compressed(std::string &out, const void *data, size_t nbytes) {
    const unsigned char *src = (const unsigned char *)data;
    size_t nblocks = nbytes == 0 ? 0 : (nbytes + BLOCK - 1) / BLOCK;
    size_t last = nbytes == 0 ? 0 : nbytes - (nblocks - 1) * BLOCK;
    if (last == BLOCK) last = 0;  // VTK convention: 0 means "last block is full"
    std::vector<uint64_t> header(3 + nblocks);
    header[0] = nblocks;
    header[1] = BLOCK;
    header[2] = last;
    std::string body;
    std::vector<unsigned char> buf(compressBound(BLOCK));
    for (size_t b = 0; b < nblocks; ++b) {
        size_t len = std::min(BLOCK, nbytes - b * BLOCK);
        uLongf clen = buf.size();
        compress2(buf.data(), &clen, src + b * BLOCK, len, 1);
        header[3 + b] = clen;
        body.append((const char *)buf.data(), clen);
    }
    out.append((const char *)header.data(), header.size() * sizeof(uint64_t));
    out.append(body);
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
    std::string appended;
    std::vector<size_t> offsets;
    offsets.push_back(appended.size());
    append_compressed(appended, points3n, sizeof(double) * 3 * (size_t)n);
    std::vector<int64_t> conn((size_t)n), offs((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        conn[i] = i;
        offs[i] = i + 1;
    }
    offsets.push_back(appended.size());
    append_compressed(appended, conn.data(), sizeof(int64_t) * (size_t)n);
    offsets.push_back(appended.size());
    append_compressed(appended, offs.data(), sizeof(int64_t) * (size_t)n);
    for (int f = 0; f < nfields; ++f) {
        offsets.push_back(appended.size());
        append_compressed(appended, data[f], sizeof(double) * (size_t)ncomps[f] * (size_t)n);
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
    fwrite(appended.data(), 1, appended.size(), fp);
    fprintf(fp, "\n  </AppendedData>\n</VTKFile>\n");
    if (fclose(fp) != 0) {
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

// ≙ save_frame!(data, sys, vars...) — IO.jl:53-75 (+ capture_frame :37-46)
extern "C" int sphmw_pvd_save_frame(sphmw_ctx *c, const char *const *fields, int32_t nfields) {
    if (!c || (nfields > 0 && !fields)) { sphmw_set_error("null argument"); return SPHMW_E_INVALID; }
    if (!c->pvd_open) { sphmw_set_error("save_frame: no pvd file is open"); return SPHMW_E_STATE; }
    const int64_t n = c->n;
    std::vector<std::vector<double>> soa(nfields + 1), aos(nfields + 1);
    std::vector<int> ncomps(nfields + 1);
    std::vector<const char *> names(nfields + 1);
    names[0] = "x";
    for (int f = 0; f < nfields; ++f) names[f + 1] = fields[f];
    for (int f = 0; f <= nfields; ++f) {
        const FieldDesc *d = sphmw_find_field(names[f]);
        if (!d) {
            sphmw_set_error("Variable %s does not exist!", names[f]);  // structs.jl:128-133
            return SPHMW_E_UNKNOWN_FIELD;
        }
        ncomps[f] = d->ncomp;
        soa[f].resize((size_t)d->ncomp * n);
        if (n) TRY(sphmw_download(c, names[f], soa[f].data(), n, d->ncomp));
        if (d->ncomp == 1) {
            aos[f].swap(soa[f]);
        } else {
            aos[f].resize((size_t)d->ncomp * n);
            for (int k = 0; k < d->ncomp; ++k)
                for (int64_t i = 0; i < n; ++i) aos[f][(size_t)i * d->ncomp + k] = soa[f][(size_t)k * n + i];
        }
    }
    std::vector<const double *> ptrs(nfields);
    for (int f = 0; f < nfields; ++f) ptrs[f] = aos[f + 1].data();
    std::string fname = "frame" + std::to_string(c->pvd_frame) + ".vtp";
    std::string path = c->pvd_dir + "/" + fname;
    TRY(sphmw_write_vtp(path.c_str(), n, aos[0].data(), nfields, names.data() + 1, ncomps.data() + 1,
                        ptrs.data()));
    c->pvd_entries.push_back(fname);
    c->pvd_frame += 1;
    return SPHMW_OK;
}

// ≙ save_pvd_file(data) — IO.jl:33-35
extern "C" int sphmw_pvd_close(sphmw_ctx *c) {
    if (!c) return SPHMW_E_INVALID;
    if (!c->pvd_open) { sphmw_set_error("save_pvd_file: no pvd file is open"); return SPHMW_E_STATE; }
    std::string path = c->pvd_dir + "/result.pvd";
    TRY(sphmw_write_pvd(path.c_str(), c->pvd_entries));
    c->pvd_open = false;
    return SPHMW_OK;
}
