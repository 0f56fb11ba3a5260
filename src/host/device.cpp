// device.cpp -- see include/analisi/device.h
#include "analisi/device.h"

#include <cstdlib>
#include <sstream>

namespace analisi_device {

Context &Context::instance() {
    static Context c;
    return c;
}

Context::Context() {
    const char *env = std::getenv("ANALISI_DEVICES");
    if (env && *env) {
        std::vector<int> ids;
        std::stringstream ss(env);
        std::string tok;
        while (std::getline(ss, tok, ',')) {
            if (!tok.empty()) ids.push_back(std::atoi(tok.c_str()));
        }
        check(agofrt_ctx_create(&ctx_, ids.data(), static_cast<int>(ids.size())), "agofrt_ctx_create");
    } else {
        check(agofrt_ctx_create(&ctx_, nullptr, -1), "agofrt_ctx_create");
    }
}

Context::~Context() {
    if (ctx_) agofrt_ctx_destroy(ctx_);
}

std::string Context::comm_unique_id() {
    char id[AGOFRT_COMM_ID_BYTES];
    check(agofrt_comm_unique_id(id), "agofrt_comm_unique_id");
    return std::string(id, AGOFRT_COMM_ID_BYTES);
}

void Context::comm_join(const std::string &id, int rank, int world) {
    if (id.size() != AGOFRT_COMM_ID_BYTES) throw std::runtime_error("comm_join: the id must be the 128 bytes of comm_unique_id()\n");
    check(agofrt_comm_join(ctx_, id.data(), rank, world), "agofrt_comm_join");
}

void PinnedBuffer::resize(size_t ndoubles) {
    if (ndoubles == n_ && ptr_) return;
    release();
    void *p = nullptr;
    check(agofrt_host_alloc(&p, ndoubles * sizeof(double)), "agofrt_host_alloc");
    ptr_ = static_cast<double *>(p);
    n_ = ndoubles;
}

void PinnedBuffer::release() {
    if (ptr_) agofrt_host_free(ptr_);
    ptr_ = nullptr;
    n_ = 0;
}

void Window::create(size_t natoms, int box_stride, const int *type_id, int ntypes, size_t max_frames) {
    release();
    check(agofrt_traj_create(&traj_, Context::instance().handle(), natoms, box_stride, type_id, ntypes, max_frames),
          "agofrt_traj_create");
    cap_ = max_frames;
    ids_set_ = false;
    ++generation_;
}

void Window::release() {
    if (traj_) agofrt_traj_destroy(traj_);
    traj_ = nullptr;
    cap_ = 0;
}

void Window::upload(size_t first, size_t n, const double *pos_aos, const double *box_internal) {
    check(agofrt_traj_upload(traj_, first, n, pos_aos, box_internal), "agofrt_traj_upload");
}

void Window::upload_wrap(size_t first, size_t n, double *pos_aos_inout, const double *box_internal) {
    check(agofrt_traj_upload_wrap(traj_, first, n, pos_aos_inout, box_internal), "agofrt_traj_upload_wrap");
}

void Window::upload_shared(size_t first, size_t n, const double *pos_aos, const double *box_internal, bool wrap, double *wrapped_out) {
    const unsigned flags = AGOFRT_UP_SHARED | (wrap ? AGOFRT_UP_WRAP : 0u) | (wrap && wrapped_out ? AGOFRT_UP_WRITEBACK : 0u);
    check(agofrt_traj_upload_ex(traj_, first, n, pos_aos, box_internal, flags, wrapped_out), "agofrt_traj_upload_ex");
}

void Window::set_rotation(size_t first, size_t n, const double *q9) {
    check(agofrt_traj_set_rotation(traj_, first, n, q9), "agofrt_traj_set_rotation");
}

void Window::get_rotation(size_t frame, double *q9) { check(agofrt_traj_get_rotation(traj_, frame, q9), "agofrt_traj_get_rotation"); }

void Window::set_ids(const int *slot_to_id, const int *slot_raw_type) {
    check(agofrt_traj_set_ids(traj_, slot_to_id, slot_raw_type), "agofrt_traj_set_ids");
    ids_set_ = true;
}

bool Window::upload_records(size_t first, size_t n, const void *const *chunk_ptr, const int *chunk_atoms, const size_t *frame_chunk,
                            const double *box_internal, bool wrap) {
    const int rc = agofrt_traj_upload_records(traj_, first, n, chunk_ptr, chunk_atoms, frame_chunk, box_internal,
                                              AGOFRT_UP_SHARED | (wrap ? AGOFRT_UP_WRAP : 0u), nullptr);
    if (rc == AGOFRT_ERR_RETYPED) return false;
    check(rc, "agofrt_traj_upload_records");
    return true;
}

void Window::download(size_t first, size_t n, double *pos_aos_out) {
    check(agofrt_traj_download(traj_, first, n, pos_aos_out), "agofrt_traj_download");
}

void pbc_wrap(double *pos_aos, size_t nframes, size_t natoms, const double *box_internal, int box_stride) {
    check(agofrt_pbc_wrap(Context::instance().handle(), pos_aos, nframes, natoms, box_internal, box_stride),
          "agofrt_pbc_wrap");
}

}  // namespace analisi_device
