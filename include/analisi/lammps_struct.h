// lammps_struct.h -- LAMMPS binary dump frames ("id type xu yu zu vx vy vz", 8 doubles per atom).
//
// Reads the three on-disk header flavours the reference reads (lib/include/lammps_struct.h:31-270):
//   * pre-2020, orthorhombic: bigint timestep, bigint natoms, int triclinic(=0), int boundary[6],
//     double box[6] (xlo xhi ylo yhi zlo zhi), int size_one, int nchunk
//   * pre-2020, triclinic: the same with double xy,xz,yz after the box
//   * 2020 format: a NEGATIVE bigint (-length of the magic string) first, the magic string, int endian,
//     int revision, then the fields above, and for revision > 1 the unit style, an optional time and
//     the column names before nchunk
// followed by nchunk chunks of {int ndoubles; double data[ndoubles]}.
// The writer emits the pre-2020 layout with one chunk (lib/src/basetrajectory.cpp:4-49).
// One flat parser instead of the reference's three structs + dispatcher.
#ifndef ANALISI_B200_LAMMPS_STRUCT_H
#define ANALISI_B200_LAMMPS_STRUCT_H

#include <cstdint>
#include <cstring>
#include <ostream>
#include <stdexcept>
#include <string>

constexpr int kLammpsDoublesPerAtom = 8;   // id type x y z vx vy vz

struct LammpsChunk {
    const char *data;   // first double of the chunk
    int natoms;
};

struct LammpsFrameHeader {
    int64_t timestep = 0;
    int64_t natoms = 0;
    int triclinic = 0;
    int boundary[6] = {0, 0, 0, 0, 0, 0};
    double box[6] = {0, 0, 0, 0, 0, 0};      // xlo xhi ylo yhi zlo zhi
    double xy_xz_yz[3] = {0, 0, 0};
    int size_one = 0;
    int nchunk = 0;
    bool format2020 = false;
    int revision = 0;

    // Parse the header that starts at `p`; returns the number of header bytes (the first chunk
    // starts at p + that).  Throws like the reference on truncated files and unsupported row sizes.
    size_t parse(const char *p, const char *end) {
        const char *q = p;
        // sizes are compared, never pointers advanced by a length that comes from the file (a corrupt magic / unit /
        // column length would overflow the pointer arithmetic); string fields longer than 64 KiB are refused
        auto need = [&](size_t n) {
            if (q > end || n > static_cast<size_t>(end - q)) throw std::runtime_error("Error: end of file reached");
        };
        auto skip_string = [&](int64_t n) {
            if (n < 0 || n > 65536) throw std::runtime_error("Error: corrupt header (length of a string field out of range)");
            need(static_cast<size_t>(n));
            q += n;
        };
        auto get = [&](auto *dst, size_t count) {
            const size_t bytes = sizeof(*dst) * count;
            need(bytes);
            std::memcpy(dst, q, bytes);
            q += bytes;
        };
        int64_t first = 0;
        get(&first, 1);
        if (first < 0) {
            format2020 = true;
            if (first == INT64_MIN) throw std::runtime_error("Error: corrupt header (length of a string field out of range)");
            skip_string(-first);
            int endian = 0;
            get(&endian, 1);
            get(&revision, 1);
            get(&timestep, 1);
        } else {
            format2020 = false;
            timestep = first;
        }
        get(&natoms, 1);
        get(&triclinic, 1);
        get(boundary, 6);
        get(box, 6);
        if (triclinic) get(xy_xz_yz, 3);
        get(&size_one, 1);
        if (format2020 && revision > 1) {
            int unit_len = 0;
            get(&unit_len, 1);
            if (unit_len > 0) skip_string(unit_len);
            char time_flag = 0;
            get(&time_flag, 1);
            if (time_flag) {
                double time = 0;
                get(&time, 1);
            }
            int columns_len = 0;
            get(&columns_len, 1);
            if (columns_len > 0) skip_string(columns_len);
        }
        get(&nchunk, 1);
        if (q >= end) throw std::runtime_error("Error: end of file reached");
        if (size_one != kLammpsDoublesPerAtom)
            throw std::runtime_error("ERROR: the binary format does not have " + std::to_string(kLammpsDoublesPerAtom) +
                                     " numbers per atom but it has " + std::to_string(size_one));
        return static_cast<size_t>(q - p);
    }

    // pre-2020 layout
    void write(std::ostream &out) const {
        out.write(reinterpret_cast<const char *>(&timestep), sizeof(timestep));
        out.write(reinterpret_cast<const char *>(&natoms), sizeof(natoms));
        out.write(reinterpret_cast<const char *>(&triclinic), sizeof(triclinic));
        out.write(reinterpret_cast<const char *>(boundary), sizeof(boundary));
        out.write(reinterpret_cast<const char *>(box), sizeof(box));
        if (triclinic) out.write(reinterpret_cast<const char *>(xy_xz_yz), sizeof(xy_xz_yz));
        out.write(reinterpret_cast<const char *>(&size_one), sizeof(size_one));
        out.write(reinterpret_cast<const char *>(&nchunk), sizeof(nchunk));
    }
};

#endif
