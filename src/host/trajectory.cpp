// trajectory.cpp -- see include/analisi/trajectory.h
#include "analisi/trajectory.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>

// Size in bytes of the frame that starts at `offset`; fills the header and, when asked, the chunk
// table (reference lib/src/trajectory.cpp:260-317).
size_t Trajectory::frame_bytes(size_t offset, LammpsFrameHeader &head, std::vector<LammpsChunk> *chunks) {
    if (offset >= fsize) throw std::runtime_error("Error: trying to read beyond the end of file");
    const char *end = file + fsize;
    size_t used = head.parse(file + offset, end);
    if (natoms != 0 && natoms != head.natoms) {
        std::stringstream ss;
        ss << "Error: at the LAMMPS timestep " << head.timestep << " the number of atoms has changed \n";
        throw std::runtime_error(ss.str());
    }
    if (head.nchunk <= 0) throw std::runtime_error("Error: the number of chunks of the timestep cannot be <= 0");
    if (chunks) chunks->clear();
    for (int c = 0; c < head.nchunk; ++c) {
        int ndouble = 0;
        if (file + offset + used + sizeof(int) > end) throw std::runtime_error("Error: end of file reached");
        std::memcpy(&ndouble, file + offset + used, sizeof(int));
        if (ndouble < 0 || ndouble % kLammpsDoublesPerAtom != 0)
            throw std::runtime_error("Number of bytes of the chunk is not a multiple of the number of atoms\n");
        used += sizeof(int);
        const size_t bytes = static_cast<size_t>(ndouble) * sizeof(double);
        if (file + offset + used + bytes > end) throw std::runtime_error("Error: end of file reached");
        if (chunks) chunks->push_back({file + offset + used, ndouble / kLammpsDoublesPerAtom});
        used += bytes;
    }
    return used;
}

Trajectory::Trajectory(std::string filename) {
    wrap_pbc = false;   // the reference's constructor resets it (lib/src/trajectory.cpp:73): callers set it afterwards
    fd = open(filename.c_str(), O_RDONLY);
    if (fd == -1) throw std::runtime_error("Error opening the trajectory \"" + filename + "\"\n");
    struct stat sb;
    if (fstat(fd, &sb) == -1) {
        close(fd);
        fd = -1;
        throw std::runtime_error("Error in finding trajectory file size \"" + filename + "\"\n");
    }
    fsize = static_cast<size_t>(sb.st_size);
    std::cerr << "Trajectory file size \"" << filename << "\": " << fsize << "\n";
    void *m = fsize ? mmap(nullptr, fsize, PROT_READ, MAP_PRIVATE, fd, 0) : MAP_FAILED;
    if (m == MAP_FAILED) {
        close(fd);
        fd = -1;
        throw std::runtime_error("mmap failed for file \"" + filename + "\".\n");
    }
    file = static_cast<char *>(m);

    // frame 0: number of atoms, cell kind, id -> slot map in order of first appearance, raw types
    LammpsFrameHeader h0;
    std::vector<LammpsChunk> chunks;
    const size_t frame0 = frame_bytes(0, h0, &chunks);
    natoms = h0.natoms;
    triclinic = h0.triclinic != 0;
    buffer_boxes_stride = triclinic ? 9 : 6;
    box_format = triclinic ? BoxFormat::Lammps_triclinic : BoxFormat::Lammps_ortho;
    raw_type.assign(natoms, -1);
    type_id.assign(natoms, -1);
    buffer_type = raw_type.data();
    buffer_type_id = type_id.data();
    slot_to_id.reserve(natoms);
    id_to_slot.reserve(static_cast<size_t>(natoms) * 2);
    for (const LammpsChunk &c : chunks) {
        for (int a = 0; a < c.natoms; ++a) {
            double rec[2];
            std::memcpy(rec, c.data + static_cast<size_t>(a) * kLammpsDoublesPerAtom * sizeof(double), sizeof(rec));
            const int id = static_cast<int>(std::round(rec[0]));
            if (id < 0) throw std::runtime_error("Error: found a negative atomic id in the trajectory");
            if (static_cast<ssize_t>(slot_to_id.size()) >= natoms && id_to_slot.find(id) == id_to_slot.end())
                throw std::runtime_error(
                    "Error: the number of atoms does not correspond to the sum of the number of atoms of each chunk\n");
            auto ins = id_to_slot.emplace(id, static_cast<int>(slot_to_id.size()));
            if (ins.second) slot_to_id.push_back(id);
            raw_type[ins.first->second] = static_cast<int>(std::round(rec[1]));
        }
    }
    if (static_cast<ssize_t>(slot_to_id.size()) != natoms)
        throw std::runtime_error("Error: the first frame does not hold one record for each of its atoms (repeated atomic ids?)\n");
    // compact ids (the usual case): a flat table instead of a hash lookup per atom and frame
    {
        int max_id = 0;
        for (int id : slot_to_id) max_id = std::max(max_id, id);
        if (static_cast<size_t>(max_id) < static_cast<size_t>(natoms) * 8 + 1024) {
            dense_slot.assign(static_cast<size_t>(max_id) + 1, -1);
            for (size_t k = 0; k < slot_to_id.size(); ++k) dense_slot[slot_to_id[k]] = static_cast<int>(k);
        }
    }
    get_ntypes();
    std::cerr << "Types of atoms (" << natoms << " atoms):";
    {
        std::vector<size_t> cnt(ntypes, 0);
        for (ssize_t i = 0; i < natoms; ++i) cnt[type_id[i]]++;
        for (ssize_t k = 0; k < ntypes; ++k) std::cerr << "  type " << types[k] << " -> index " << k << " (" << cnt[k] << ")";
        std::cerr << "\n";
    }
    // estimate of the number of frames from the size of the first one (reference :114)
    n_timesteps = static_cast<ssize_t>(fsize / frame0);
    if (const char *e = std::getenv("ANALISI_PREFETCH")) prefetch_enabled = std::atoi(e) != 0;
    offsets.assign(static_cast<size_t>(n_timesteps) + 1, 0);
    lammps_steps.assign(static_cast<size_t>(n_timesteps) + 1, 0);
    indexed_upto = 0;
    lammps_steps[0] = h0.timestep;
}

Trajectory::~Trajectory() {
    cancel_prefetch();
    if (file) munmap(file, fsize);
    if (fd != -1) close(fd);
    buffer_positions = buffer_velocity = buffer_boxes = nullptr;
}

Trajectory::Errori Trajectory::set_data_access_block_size(const size_t &n) {
    if (!file) {
        std::cerr << "mmap not correctly initialized!\n";
        return non_inizializzato;
    }
    // (a set_load_velocities() toggle since the last call changes which buffers exist: then go on and rebuild them)
    if (static_cast<size_t>(loaded_timesteps) == n && window_capacity == n && load_velocities == (buffer_velocity != nullptr)) return Ok;
    cancel_prefetch();
    window_loaded = false;
    pos_buf.resize(n * natoms * 3);
    buffer_positions = pos_buf.data();
    if (load_velocities) {
        vel_buf.resize(n * natoms * 3);
        buffer_velocity = vel_buf.data();
        cm_pos.assign(n * ntypes * 3, 0.0);
        cm_vel.assign(n * ntypes * 3, 0.0);
    } else {
        vel_buf.release();
        buffer_velocity = nullptr;
        cm_pos.clear();
        cm_vel.clear();
    }
    boxes.assign(n * buffer_boxes_stride, 0.0);
    buffer_boxes = boxes.data();
    window_capacity = n;
    loaded_timesteps = static_cast<ssize_t>(n);
    mark_window_changed();
    return Ok;
}

// lazily index frame offsets up to frame `upto` by walking the headers (reference :481-494)
void Trajectory::ensure_indexed(size_t upto) {
    if (upto + 1 > offsets.size()) {
        offsets.resize(upto + 2, 0);
        lammps_steps.resize(upto + 2, 0);
    }
    while (indexed_upto < upto) {
        LammpsFrameHeader h;
        const size_t sz = frame_bytes(offsets[indexed_upto], h, nullptr);
        lammps_steps[indexed_upto] = h.timestep;
        offsets[indexed_upto + 1] = offsets[indexed_upto] + sz;
        ++indexed_upto;
    }
}

void Trajectory::index_all() {
    if (n_timesteps > 0) ensure_indexed(static_cast<size_t>(n_timesteps) - 1);
    if (n_timesteps > 0) {
        LammpsFrameHeader h;
        frame_bytes(offsets[n_timesteps - 1], h, nullptr);
        lammps_steps[n_timesteps - 1] = h.timestep;
    }
}

// one frame of the file -> window slot: box row (internal format), per-atom scatter by id
// (reference :593-662)
void Trajectory::read_frame_into_slot(size_t frame, size_t slot) {
    read_frame_to(frame, buffer_positions + slot * natoms * 3, buffer_boxes + slot * buffer_boxes_stride,
                  buffer_velocity ? buffer_velocity + slot * natoms * 3 : nullptr,
                  buffer_velocity ? cm_pos.data() + slot * ntypes * 3 : nullptr,
                  buffer_velocity ? cm_vel.data() + slot * ntypes * 3 : nullptr);
}

void Trajectory::read_frame_to(size_t frame, double *P, double *b, double *V, double *cp, double *cv) {
    // the caller has indexed the file up to `frame` (ensure_indexed): this function only reads, so
    // several frames can be read by several threads at once
    LammpsFrameHeader h;
    std::vector<LammpsChunk> chunks;
    frame_bytes(offsets[frame], h, &chunks);
    lammps_steps[frame] = h.timestep;
    if ((h.triclinic != 0) != triclinic) throw std::runtime_error("Error: the cell kind (triclinic flag) changes along the trajectory\n");
    std::memcpy(b, h.box, 6 * sizeof(double));
    if (triclinic) std::memcpy(b + 6, h.xy_xz_yz, 3 * sizeof(double));
    lammps_to_internal(b);
    std::vector<size_t> cnt;
    if (V) {
        cnt.assign(ntypes, 0);
        std::fill(cp, cp + ntypes * 3, 0.0);
        std::fill(cv, cv + ntypes * 3, 0.0);
    }
    for (const LammpsChunk &c : chunks) {
        const char *p = c.data;
        for (int a = 0; a < c.natoms; ++a, p += kLammpsDoublesPerAtom * sizeof(double)) {
            double rec[kLammpsDoublesPerAtom];
            std::memcpy(rec, p, sizeof(rec));
            const int id = static_cast<int>(std::round(rec[0]));
            int found = -1;
            if (!dense_slot.empty()) {
                if (id >= 0 && static_cast<size_t>(id) < dense_slot.size()) found = dense_slot[id];
            } else {
                auto it = id_to_slot.find(id);
                if (it != id_to_slot.end()) found = it->second;
            }
            if (found < 0) throw std::out_of_range("atom id that was not in the first frame");
            const size_t s = static_cast<size_t>(found);
            P[3 * s] = rec[2];
            P[3 * s + 1] = rec[3];
            P[3 * s + 2] = rec[4];
            const int tipo = static_cast<int>(std::round(rec[1]));
            if (raw_type[s] != tipo) {
                // frames are read by several threads (and by the read-ahead thread): one at a time here
                std::lock_guard<std::mutex> lock(type_change_mutex);
                if (raw_type[s] != tipo) {
                    std::cerr << "WARNING: atomic type for atom with id " << s << " is changing from " << raw_type[s] << " to "
                              << tipo << " !\n";
                    raw_type[s] = tipo;
                }
            }
            if (V) {
                V[3 * s] = rec[5];
                V[3 * s + 1] = rec[6];
                V[3 * s + 2] = rec[7];
                // running per-type mean, same update as the reference (:648-657)
                const int k = type_id[s];
                cnt[k]++;
                for (int d = 0; d < 3; ++d) {
                    cp[3 * k + d] += (P[3 * s + d] - cp[3 * k + d]) / cnt[k];
                    cv[3 * k + d] += (V[3 * s + d] - cv[3 * k + d]) / cnt[k];
                }
            }
        }
    }
}

// Frames [first, last) of the file -> rows (frame - origin) of a window buffer, by several host threads.
// Frames are independent (each fills its own row).  The per-type centres of mass are running means in
// file order inside ONE frame, so they do not constrain the split either.
void Trajectory::read_frames(size_t first, size_t last, size_t origin, double *P0, double *B0, bool own_window) {
    const size_t nfr = last - first;
    auto one = [&](size_t f) {
        if (own_window)
            read_frame_into_slot(f, f - origin);
        else
            read_frame_to(f, P0 + (f - origin) * natoms * 3, B0 + (f - origin) * buffer_boxes_stride, nullptr, nullptr, nullptr);
    };
    size_t nth = std::min<size_t>({nfr, std::max(1u, std::thread::hardware_concurrency()), 32});
    if (static_cast<size_t>(natoms) * nfr < 200000) nth = 1;
    if (nth <= 1) {
        for (size_t f = first; f < last; ++f) one(f);
        return;
    }
    std::vector<std::thread> pool;
    std::vector<std::exception_ptr> errors(nth);
    std::atomic<size_t> next{first};
    for (size_t t = 0; t < nth; ++t)
        pool.emplace_back([&, t]() {
            try {
                for (size_t f = next.fetch_add(1); f < last; f = next.fetch_add(1)) one(f);
            } catch (...) {
                errors[t] = std::current_exception();
            }
        });
    for (std::thread &th : pool) th.join();
    for (const std::exception_ptr &e : errors)
        if (e) std::rethrow_exception(e);   // first failure, on the caller's thread (as the reference does)
}

bool Trajectory::device_parse_possible() const {
    if (load_velocities || dense_slot.empty() || natoms == 0) return false;
    // ANALISI_DEVICE_PARSE=1 / 0 forces it on / off; by default frames of at least 4096 atoms go that way (measured on
    // the 56-atom test trajectory the host threads parse a 757-frame window faster than the 1.8 ms a device round
    // trip costs; from a few thousand atoms on the id scatter is what the GPU does better)
    static const int forced = [] {
        const char *e = std::getenv("ANALISI_DEVICE_PARSE");
        return e ? (std::atoi(e) != 0 ? 1 : 0) : -1;
    }();
    return forced >= 0 ? forced == 1 : natoms >= 4096;
}

// Headers of frames [first, first + n) -- box rows (internal format) into B0, LAMMPS timesteps -- and the table of their
// chunks of raw records, which stay where they are in the mapping.  The caller has indexed the file up to the last
// frame (ensure_indexed): this function only reads, like read_frame_to.
void Trajectory::gather_records(size_t first, size_t n, RecordTable &tab, double *B0) {
    tab.ptr.clear();
    tab.atoms.clear();
    tab.frame_chunk.assign(1, 0);
    std::vector<LammpsChunk> chunks;
    for (size_t f = first; f < first + n; ++f) {
        LammpsFrameHeader h;
        frame_bytes(offsets[f], h, &chunks);
        lammps_steps[f] = h.timestep;
        if ((h.triclinic != 0) != triclinic) throw std::runtime_error("Error: the cell kind (triclinic flag) changes along the trajectory\n");
        double *b = B0 + (f - first) * buffer_boxes_stride;
        std::memcpy(b, h.box, 6 * sizeof(double));
        if (triclinic) std::memcpy(b + 6, h.xy_xz_yz, 3 * sizeof(double));
        lammps_to_internal(b);
        for (const LammpsChunk &c : chunks) {
            tab.ptr.push_back(c.data);
            tab.atoms.push_back(c.natoms);
        }
        tab.frame_chunk.push_back(tab.ptr.size());
    }
}

// Read-ahead: callers that walk the file in equal steps (BlockAverageG: block after block) find the next
// window already parsed into a second page-locked buffer, read by a background thread while the GPUs were
// busy with the current block.  Only the host part runs ahead (parse + scatter); wrap and upload stay on the
// caller's thread, as the C ABI wants one host thread per context.
void Trajectory::cancel_prefetch() {
    if (prefetch_thread.joinable()) prefetch_thread.join();
    prefetch_valid = false;
    prefetch_error = nullptr;
}

void Trajectory::start_prefetch(size_t target) {
    if (!prefetch_enabled || load_velocities) return;
    const size_t W = static_cast<size_t>(loaded_timesteps);
    if (target + W > static_cast<size_t>(n_timesteps)) return;
    try {
        ensure_indexed(target + W - 1);
        LammpsFrameHeader h;
        frame_bytes(offsets[target + W - 1], h, nullptr);   // the last frame must be complete
    } catch (const std::exception &) {
        return;   // the estimate of the number of frames was too generous: nothing to read ahead
    }
    pos_alt.resize(W * natoms * 3);
    boxes_alt.assign(W * buffer_boxes_stride, 0.0);
    prefetch_target = target;
    prefetch_valid = true;
    prefetch_error = nullptr;
    prefetch_uploaded = false;
    const bool to_device = device_active();   // decided on the caller's thread
    const bool records = to_device && device_parse_possible();
    prefetch_records = false;
    prefetch_thread = std::thread([this, target, W, to_device, records]() {
        try {
            if (records) {
                // the raw records go to the second device window and are parsed, wrapped and laid out there, under
                // the pair kernels of the block the caller is computing; the host copy follows on demand
                RecordTable tab;
                gather_records(target, W, tab, boxes_alt.data());
                if (upload_next_window_records(target, W, tab.ptr.data(), tab.atoms.data(), tab.frame_chunk.data(), boxes_alt.data(),
                                               wrap_pbc, slot_to_id.data(), raw_type.data())) {
                    prefetch_uploaded = true;
                    prefetch_records = true;
                    return;
                }
                // an atom changed type in that window: the host parser reads it (and warns, as the reference does)
            }
            read_frames(target, target + W, target, pos_alt.data(), boxes_alt.data(), false);
            if (to_device) {
                // wrap on the device, lay the window out there, bring the wrapped frames back: all of it under
                // the pair kernels of the block the caller is computing
                upload_next_window(target, W, pos_alt.data(), boxes_alt.data(), wrap_pbc);
                prefetch_uploaded = true;
            }
        } catch (...) {
            prefetch_error = std::current_exception();
        }
    });
}

Trajectory::Errori Trajectory::set_access_at(const size_t &timestep) {
    const auto t_begin = std::chrono::steady_clock::now();
    if (!file) {
        std::cerr << "mmap not initialized correctly!\n";
        return non_inizializzato;
    }
    if (loaded_timesteps <= 0 || !buffer_positions) throw std::runtime_error("set_data_access_block_size must be called first\n");
    if (timestep == static_cast<size_t>(current_timestep) && window_loaded) return Ok;
    const size_t W = static_cast<size_t>(loaded_timesteps);
    const bool had_window = window_loaded;
    const size_t previous = static_cast<size_t>(current_timestep);
    auto finish = [&](const char *how, bool on_device, bool host_copy_is_current = true, bool uploaded_here = false) {
        current_timestep = static_cast<ssize_t>(timestep);
        window_loaded = true;
        if (!uploaded_here) mark_window_changed();   // (upload_records_now has already matched the epochs)
        if (on_device) adopt_next_window(host_copy_is_current);
        // equal steps forward (or the stride the block loop announced): read the next window ahead
        if (had_window && timestep > previous)
            start_prefetch(timestep + (timestep - previous));
        else if (stride_hint > 0)
            start_prefetch(timestep + stride_hint);
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
        std::cerr << "Reading time: " << dt << "s" << how << ".\n";
        return Ok;
    };
    if (prefetch_thread.joinable() || prefetch_valid) {
        if (prefetch_thread.joinable()) prefetch_thread.join();
        const bool hit = prefetch_valid && !prefetch_error && prefetch_target == timestep && !load_velocities &&
                         pos_alt.size() == W * static_cast<size_t>(natoms) * 3;
        prefetch_valid = false;
        prefetch_error = nullptr;
        if (hit) {
            pos_buf.swap(pos_alt);
            boxes.swap(boxes_alt);
            buffer_positions = pos_buf.data();
            buffer_boxes = boxes.data();
            window_loaded = false;
            if (prefetch_uploaded && prefetch_records) return finish(" (read ahead: parsed on the GPUs)", true, false);
            if (prefetch_uploaded) return finish(" (read ahead, already on the GPUs)", true);
            if (wrap_pbc) pbc_wrap_frames(0, W);
            return finish(" (read ahead)", false);
        }
    }

    // Device-side ingest: the records of the window go to the GPUs as they are and are parsed there
    if (device_parse_possible()) {
        ensure_indexed(timestep + W - 1);
        RecordTable tab;
        gather_records(timestep, W, tab, buffer_boxes);
        madvise(file, offsets[timestep] & ~static_cast<size_t>(sysconf(_SC_PAGESIZE) - 1), MADV_DONTNEED);
        current_timestep = static_cast<ssize_t>(timestep);   // upload_records_now uploads [current_timestep, +loaded_timesteps)
        if (upload_records_now(tab.ptr.data(), tab.atoms.data(), tab.frame_chunk.data(), wrap_pbc, slot_to_id.data(), raw_type.data()))
            return finish(" (parsed on the GPUs)", false, false, true);
        current_timestep = static_cast<ssize_t>(previous);    // an atom changed type: the host parser below reads the window
        window_loaded = false;
    }

    // overlap with what is already in the window: move it instead of reading it again (reference :542-586)
    size_t read_begin = timestep, read_end = timestep + W;   // frames to read from the file
    if (window_loaded) {
        const ssize_t shift = static_cast<ssize_t>(current_timestep) - static_cast<ssize_t>(timestep);  // old slot + shift = new slot
        const size_t ashift = static_cast<size_t>(shift < 0 ? -shift : shift);
        if (ashift < W) {
            const size_t keep = W - ashift;
            const size_t src = shift < 0 ? ashift : 0, dst = shift < 0 ? 0 : ashift;
            auto move_rows = [&](double *base, size_t row) {
                if (base) std::memmove(base + dst * row, base + src * row, keep * row * sizeof(double));
            };
            move_rows(buffer_positions, natoms * 3);
            move_rows(buffer_velocity, natoms * 3);
            move_rows(buffer_boxes, buffer_boxes_stride);
            if (buffer_velocity) {
                move_rows(cm_pos.data(), ntypes * 3);
                move_rows(cm_vel.data(), ntypes * 3);
            }
            if (shift < 0) {
                read_begin = current_timestep + W;   // the tail is new
            } else {
                read_end = current_timestep;         // the head is new
            }
        }
    }
    window_loaded = false;
    ensure_indexed(timestep);
    // tell the kernel what we are done with and what comes next (reference :497-528)
    madvise(file, offsets[timestep] & ~static_cast<size_t>(sysconf(_SC_PAGESIZE) - 1), MADV_DONTNEED);
    if (read_end > read_begin) {
        ensure_indexed(read_end - 1);
        read_frames(read_begin, read_end, timestep, buffer_positions, buffer_boxes, true);
    }
    if (wrap_pbc && read_end > read_begin) pbc_wrap_frames(static_cast<ssize_t>(read_begin - timestep), read_end - read_begin);
    return finish("", false);
}

int64_t Trajectory::get_timestep_lammps(size_t timestep) {
    if (timestep < static_cast<size_t>(n_timesteps) && timestep < lammps_steps.size()) return lammps_steps[timestep];
    std::stringstream ss;
    ss << "Error: requested to read a timestep that probably is beyond the end of the file (" << timestep << ", "
       << n_timesteps << " letti)\n";
    throw std::runtime_error(ss.str());
}

int *Trajectory::get_lammps_id() {
    int *out = new int[natoms];
    for (ssize_t i = 0; i < natoms; ++i) out[i] = slot_to_id[i];
    return out;
}

int *Trajectory::get_lammps_type() {
    int *out = new int[natoms];
    for (ssize_t i = 0; i < natoms; ++i) out[i] = raw_type[i];
    return out;
}
