// fused_step.cu -- one reconstruction train step of the occupancy (cuda_ray) path as a fixed sequence of launches on
// one stream, driven by a plan of device pointers.
//
// Replaces, for the training hot loop, the Python glue of the reference between its native ops:
//   NeRFRenderer.run_cuda            nerf/renderer.py:597-650   near/far -> march -> field -> composite
//   NeRFNetwork.forward              nerf/network_grid.py:159-177
//   _grid_encode / _composite autograd Functions  gridencoder/grid.py:27-95, raymarching/raymarching.py:239-292
//   MSE loss + scaler + optimiser    nerf/utils_init_nerf.py:224-234, 612-629
//
// Nothing here synchronises with the host or allocates: the sample count of the step stays on the device
// (`m_eff`), every buffer has a fixed capacity (`M_cap` rows) and the kernels bound their work by the device-side
// count, so the whole step can be captured once into a CUDA graph and replayed (the reference blocks on
// counter[0].item() every step, raymarching.py:225).  Rays whose samples do not fit in M_cap rows are dropped
// exactly as the reference drops them when its mean_count budget overflows (raymarching.cu:415-416); the caller
// watches counter[0] against M_cap and grows the buffers.
#include "common.cuh"
#include "wgrad_reduce.cuh"

namespace {
struct StageTimer {
    cudaEvent_t fb[NB200_FB_STAGES + 1];
    cudaEvent_t up[NB200_UP_STAGES + 1];
    bool fb_done, up_done;
};
inline void tick(cudaEvent_t *ev, int i, cudaStream_t st) { if (ev) cudaEventRecord(ev[i], st); }

// grid encode + field network forward of the step: two launches, or ONE (NB200_PLAN_FUSED_FORWARD: the gather runs in
// producer warps of the field kernel, csrc/field_fused.cu; x_en is still written once, for the backward pass)
int fs_encode_field_forward(const nb200_train_plan *p, void *stream, cudaEvent_t *ev, int *k) {
    cudaStream_t st = nb_stream(stream);
    int rc;
    if (p->flags & NB200_PLAN_FUSED_FORWARD) {
        if (k) tick(ev, (*k)++, st);            // the encode stage is empty
        if ((rc = nb200_field_fused_forward(p->xyzs, p->dirs, p->bound, p->table, p->offsets, p->L, p->S, p->base_res, p->gridtype,
                                            0, 0, p->w_fwd, p->sigma, p->sigma_arg, p->rgba, p->x_en, p->act, p->M_cap, p->m_eff,
                                            stream))) return rc;
        if (k) tick(ev, (*k)++, st);
        return 0;
    }
    if ((rc = nb200_fs_encode_forward(p->xyzs, p->bound, p->table, p->offsets, p->x_en, p->M_cap, p->L, p->S, p->base_res,
                                      p->gridtype, 0, 0, p->m_eff, stream))) return rc;
    if (k) tick(ev, (*k)++, st);
    if ((rc = nb200_field_forward(p->x_en, p->xyzs, p->dirs, p->w_fwd, p->sigma, p->sigma_arg, p->rgba, p->act, p->M_cap,
                                  p->m_eff, stream))) return rc;
    if (k) tick(ev, (*k)++, st);
    return 0;
}
}  // namespace

extern "C" {

int nb200_stage_timer_create(void **timer) {
    if (!timer) return NB200_E_BAD_ARG;
    StageTimer *t = new StageTimer();
    t->fb_done = t->up_done = false;
    for (auto &e : t->fb) { cudaError_t rc = cudaEventCreate(&e); if (rc != cudaSuccess) return (int)rc; }
    for (auto &e : t->up) { cudaError_t rc = cudaEventCreate(&e); if (rc != cudaSuccess) return (int)rc; }
    *timer = t;
    return 0;
}

int nb200_stage_timer_destroy(void *timer) {
    StageTimer *t = (StageTimer *)timer;
    if (!t) return 0;
    for (auto &e : t->fb) cudaEventDestroy(e);
    for (auto &e : t->up) cudaEventDestroy(e);
    delete t;
    return 0;
}

int nb200_stage_timer_read(void *timer, float *out_us) {
    StageTimer *t = (StageTimer *)timer;
    if (!t || !out_us || !t->fb_done || !t->up_done) return NB200_E_BAD_ARG;
    cudaError_t rc = cudaEventSynchronize(t->up[NB200_UP_STAGES]);
    if (rc != cudaSuccess) return (int)rc;
    for (int i = 0; i < NB200_FB_STAGES; i++) {
        float ms = 0; cudaEventElapsedTime(&ms, t->fb[i], t->fb[i + 1]); out_us[i] = ms * 1e3f;
    }
    for (int i = 0; i < NB200_UP_STAGES; i++) {
        float ms = 0; cudaEventElapsedTime(&ms, t->up[i], t->up[i + 1]); out_us[NB200_FB_STAGES + i] = ms * 1e3f;
    }
    return 0;
}


// phases: NB200_PHASE_MARCH = near/far + march (reads only the rays, the noises and the occupancy bit field -- nothing the
// optimiser writes, so it may run concurrently with the previous step's nb200_train_update), NB200_PHASE_REST = encode ..
// encode^T.  nb200_train_forward_backward = both.
int nb200_train_phase(const nb200_train_plan *p, int phases, void *stream) {
    if (!p) return NB200_E_BAD_ARG;
    cudaStream_t st = nb_stream(stream);
    int rc;
    cudaError_t e;
    StageTimer *tm = (StageTimer *)p->timer;
    cudaEvent_t *ev = tm ? tm->fb : nullptr;
    int k = 0;
    if (!(phases & NB200_PHASE_MARCH)) { k = 2; goto rest; }
    if ((e = cudaMemsetAsync(p->counter, 0, 2 * sizeof(int32_t), st)) != cudaSuccess) return (int)e;
    tick(ev, k++, st);
    if ((rc = nb200_fs_march_count(p->rays_o, p->rays_d, p->bitfield, p->bound, p->dt_gamma, p->max_steps, p->N, p->C,
                                   p->H, p->M_cap, p->nears, p->fars, p->noises, p->rays, p->counter, p->m_eff,
                                   p->scratch, p->aabb, p->min_near, stream))) return rc;
    tick(ev, k++, st);
    if ((rc = nb200_fs_march_write(p->rays_o, p->rays_d, p->bitfield, p->bound, p->dt_gamma, p->max_steps, p->N, p->C,
                                   p->H, p->M_cap, p->nears, p->fars, p->noises, p->rays, p->xyzs, p->dirs, p->deltas,
                                   p->m_eff, p->scratch, stream))) return rc;
rest:
    if (phases & NB200_PHASE_REST_B) {      // split step, second scatter launch: the coarse levels
        if ((rc = nb200_fs_encode_backward_levels(p->d_x_en, p->xyzs, p->bound, p->offsets, p->g_table, p->M_cap, p->L, p->S,
                                                  p->base_res, p->gridtype, 0, 0, p->m_eff, 0, p->split_level, stream))) return rc;
        return 0;
    }
    if (!(phases & (NB200_PHASE_REST | NB200_PHASE_REST_A))) return 0;
    if ((e = cudaMemsetAsync(p->loss, 0, sizeof(float), st)) != cudaSuccess) return (int)e;
    tick(ev, k++, st);
    if ((rc = fs_encode_field_forward(p, stream, ev, &k))) return rc;
    if (!(p->flags & NB200_PLAN_SPLIT_COMPOSITE) && p->target) {
        // compositing forward + MSE + compositing backward in one launch (the stage timer then books it all under the forward)
        if ((rc = nb200_fs_composite_fused(p->sigma, p->rgba, p->deltas, p->rays, p->M_cap, p->N, p->T_thresh, p->weights_sum,
                                           p->depth, p->image, p->target, p->inv_n_total, p->loss_scale, p->loss, p->g_image,
                                           p->target_mask, p->mask_weight, p->render_mask, p->g_render_mask,
                                           (const float *)p->scaler, p->g_weights_sum, p->d_sigma, p->d_rgba, stream))) return rc;
        tick(ev, k++, st);
        tick(ev, k++, st);
    } else {
        if ((rc = nb200_fs_composite_forward(p->sigma, p->rgba, p->deltas, p->rays, p->M_cap, p->N, p->T_thresh,
                                             p->weights_sum, p->depth, p->image, p->target, p->inv_n_total, p->loss_scale,
                                             p->loss, p->g_image, p->target_mask, p->mask_weight, p->render_mask,
                                             p->g_render_mask, (const float *)p->scaler, stream))) return rc;
        tick(ev, k++, st);
        if ((rc = nb200_fs_composite_backward(p->g_weights_sum, p->g_image, p->sigma, p->rgba, p->deltas, p->rays,
                                              p->weights_sum, p->image, p->M_cap, p->N, p->T_thresh, p->d_sigma, p->d_rgba,
                                              p->target_mask ? p->g_render_mask : nullptr, p->render_mask, stream))) return rc;
        tick(ev, k++, st);
    }
    {
        // the slab reduction of the field backward (the MLPs' weight gradients) depends on nothing but the slabs and only the
        // optimiser depends on it: its blocks ride at the tail of the table-scatter launch instead of being a launch of their
        // own between the two kernels (10 us of launch gap + run time off the critical path at configs[1])
        const bool ride = p->wg_scratch != nullptr;
        if ((rc = nb_field_backward_launch(p->d_sigma, p->d_rgba, p->sigma_arg, p->rgba, p->x_en, p->dirs, p->act, p->w_bwd,
                                           p->d_x_en, p->g_trunk, p->g_density, p->g_rgb, p->M_cap, p->m_eff, p->wg_scratch,
                                           p->scaler, !ride, stream))) return rc;
        tick(ev, k++, st);
        const NbWgradRed red{p->wg_scratch, p->m_eff, p->g_trunk, p->g_density, p->g_rgb, p->scaler,
                             nb_field_backward_grid(p->M_cap), p->M_cap, ride ? nb_wgrad_reduce_blocks() : 0u};
        if ((rc = nb_fs_encode_backward_red(p->d_x_en, p->xyzs, p->bound, p->offsets, p->g_table, p->M_cap, p->L, p->S, p->base_res,
                                            p->gridtype, 0, 0, p->m_eff, (phases & NB200_PHASE_REST_A) ? p->split_level : 0,
                                            p->L, &red, stream))) return rc;
        tick(ev, k++, st);
    }
    // the step's sample count, parked where the (possibly concurrent: pipelined update next to the NEXT step's march, which
    // resets the counter) update reads it for the status word of the loss scaler
    if (p->scaler && (e = cudaMemcpyAsync(p->counter + 6, p->counter, sizeof(int32_t), cudaMemcpyDeviceToDevice, st)) != cudaSuccess)
        return (int)e;
    if (tm) tm->fb_done = true;
    return 0;
}

uint32_t nb200_lgie_plan_bytes(void) { return (uint32_t)sizeof(nb200_lgie_plan); }

int nb200_train_lgie_forward(const nb200_train_plan *p, const nb200_lgie_plan *g, void *stream) {
    if (!p || !g || !g->weights_sum || !g->depth || !g->image || !g->render_mask) return NB200_E_BAD_ARG;
    int rc;
    if ((rc = nb200_train_phase(p, NB200_PHASE_MARCH, stream))) return rc;
    if ((rc = fs_encode_field_forward(p, stream, nullptr, nullptr))) return rc;
    const size_t N = p->N;
    for (int v = 0; v < 3; v++)
        if ((rc = nb200_fs_composite_lgie_forward(v, p->sigma, p->rgba, p->deltas, p->rays, p->M_cap, p->N, p->T_thresh,
                                                  g->conf_thr, g->soft_mask, g->weights_sum + v * N, g->depth + v * N,
                                                  g->image + v * N * 3, g->render_mask + v * N, stream))) return rc;
    return 0;
}

int nb200_train_lgie_backward(const nb200_train_plan *p, const nb200_lgie_plan *g, void *stream) {
    if (!p || !g || !g->g_weights_sum || !g->g_image || !g->g_render_mask) return NB200_E_BAD_ARG;
    int rc;
    const size_t N = p->N;
    for (int v = 0; v < 3; v++)
        if ((rc = nb200_fs_composite_lgie_backward(v, g->g_weights_sum + v * N, g->g_image + v * N * 3, g->g_render_mask + v * N,
                                                   p->sigma, p->rgba, p->deltas, p->rays, g->weights_sum + v * N,
                                                   g->image + v * N * 3, g->render_mask + v * N, p->M_cap, p->N, p->T_thresh,
                                                   g->conf_thr, g->soft_mask, g->detach_bg, g->detach_mask_from_field,
                                                   p->d_sigma, p->d_rgba, stream))) return rc;
    if ((rc = nb200_field_backward(p->d_sigma, p->d_rgba, p->sigma_arg, p->rgba, p->x_en, p->dirs, p->act, p->w_bwd,
                                   p->d_x_en, p->g_trunk, p->g_density, p->g_rgb, p->M_cap, p->m_eff, p->wg_scratch, p->scaler, stream))) return rc;
    if ((rc = nb200_fs_encode_backward(p->d_x_en, p->xyzs, p->bound, p->offsets, p->g_table, p->M_cap, p->L, p->S,
                                       p->base_res, p->gridtype, 0, 0, p->m_eff, stream))) return rc;
    if (p->scaler) {
        cudaError_t e = cudaMemcpyAsync(p->counter + 6, p->counter, sizeof(int32_t), cudaMemcpyDeviceToDevice, nb_stream(stream));
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}

int nb200_train_forward_backward(const nb200_train_plan *p, void *stream) {
    return nb200_train_phase(p, NB200_PHASE_MARCH | NB200_PHASE_REST, stream);
}

// dynamic loss scale: hyper-parameters for step + 1 without committing it (the sweep is skipped on a non-finite gradient and
// the commit rides with the weight re-pack); constant scale: the classic hyper kernel that advances the step itself
int nb200_train_update_hyper(const nb200_train_plan *p, int peer, void *stream) {
    if (!p) return NB200_E_BAD_ARG;
    if (p->scaler) return nb200_adam_hyper_scaled(p->step, p->sched, p->hyper, p->scaler, peer ? 0 : 1, p->counter + 6, stream);
    return nb200_adam_hyper(p->step, p->sched, p->hyper, stream);
}

int nb200_train_update(const nb200_train_plan *p, void *stream) {
    if (!p) return NB200_E_BAD_ARG;
    int rc;
    cudaStream_t st = nb_stream(stream);
    StageTimer *tm = (StageTimer *)p->timer;
    cudaEvent_t *ev = tm ? tm->up : nullptr;
    tick(ev, 0, st);
    if (!(p->flags & NB200_PLAN_HYPER_DONE) && (rc = nb200_train_update_hyper(p, 0, stream))) return rc;
    if ((rc = nb200_fused_adam_cfg(p->params_flat, p->grads_flat, p->exp_avg, p->exp_avg_sq, p->n_params, p->n_table_params,
                                   p->hyper, 1, p->adam_grid, p->adam_threads, p->adam_unroll, stream))) return rc;
    tick(ev, 1, st);
    if (p->scaler) {        // weight re-pack + GradScaler.update() / step count in one launch
        if ((rc = nb200_field_pack_weights_commit(p->trunk, p->density, p->rgb, p->w_fwd, p->w_bwd, p->step, p->scaler, nullptr, 1,
                                                  p->counter + 5, stream))) return rc;
    } else if ((rc = nb200_field_pack_weights(p->trunk, p->density, p->rgb, p->w_fwd, p->w_bwd, stream))) return rc;
    tick(ev, 2, st);
    if (tm) tm->up_done = true;
    return 0;
}

int nb200_train_update_peer(const nb200_train_plan *p, const nb200_peer_plan *peer, void *stream) {
    if (!p || !peer) return NB200_E_BAD_ARG;
    if (peer->n != p->n_params || peer->split != p->n_table_params || peer->params[peer->rank] != p->params_flat ||
        peer->grads[peer->rank] != p->grads_flat) return NB200_E_BAD_ARG;
    int rc;
    cudaStream_t st = nb_stream(stream);
    StageTimer *tm = (StageTimer *)p->timer;
    cudaEvent_t *ev = tm ? tm->up : nullptr;
    tick(ev, 0, st);
    const bool scaled = p->scaler != nullptr;
    if (scaled && peer->scalers[peer->rank] != p->scaler) return NB200_E_BAD_ARG;   // the peers must be able to read the flags
    if (!(p->flags & NB200_PLAN_HYPER_DONE) && (rc = nb200_train_update_hyper(p, 1, stream))) return rc;
    if ((rc = nb200_peer_reduce_adam_bcast(peer, stream))) return rc;
    tick(ev, 1, st);
    if (scaled) {
        if ((rc = nb200_field_pack_weights_commit(p->trunk, p->density, p->rgb, p->w_fwd, p->w_bwd, p->step, p->scaler,
                                                  peer->scalers, peer->world, p->counter + 5, stream))) return rc;
    } else if ((rc = nb200_field_pack_weights(p->trunk, p->density, p->rgb, p->w_fwd, p->w_bwd, stream))) return rc;
    tick(ev, 2, st);
    if (tm) tm->up_done = true;
    return 0;
}

int nb200_train_update_peer_part(const nb200_train_plan *p, const nb200_peer_plan *peer, int part, void *stream) {
    if (!p || !peer || (part != 1 && part != 2) || p->split_level == 0 || p->split_elem == 0 || p->split_elem >= p->n_table_params ||
        (p->split_elem & 3u)) return NB200_E_BAD_ARG;
    if (peer->n != p->n_params || peer->split != p->n_table_params || peer->params[peer->rank] != p->params_flat ||
        peer->grads[peer->rank] != p->grads_flat) return NB200_E_BAD_ARG;
    const bool scaled = p->scaler != nullptr;
    if (scaled && peer->scalers[peer->rank] != p->scaler) return NB200_E_BAD_ARG;
    int rc;
    const uint64_t e0 = part == 1 ? p->split_elem : 0, e1 = part == 1 ? p->n_params : p->split_elem;
    if (part == 1 && !(p->flags & NB200_PLAN_HYPER_DONE) && (rc = nb200_train_update_hyper(p, 1, stream))) return rc;
    nb200_peer_plan sub = *peer;            // the same kernel on a sub-range: rank r owns the r-th 1/world of [e0, e1)
    for (uint32_t q = 0; q < peer->world; q++) { sub.params[q] += e0; sub.grads[q] += e0; }
    sub.exp_avg += e0; sub.exp_avg_sq += e0;
    if (sub.mc_params) { sub.mc_params += e0; sub.mc_grads += e0; }
    sub.n = e1 - e0;
    sub.split = p->n_table_params > e0 ? (p->n_table_params - e0 < sub.n ? p->n_table_params - e0 : sub.n) : 0;
    if ((rc = nb200_peer_reduce_adam_bcast(&sub, stream))) return rc;
    if (part == 2) {
        if (scaled) {
            if ((rc = nb200_field_pack_weights_commit(p->trunk, p->density, p->rgb, p->w_fwd, p->w_bwd, p->step, p->scaler,
                                                      peer->scalers, peer->world, p->counter + 5, stream))) return rc;
        } else if ((rc = nb200_field_pack_weights(p->trunk, p->density, p->rgb, p->w_fwd, p->w_bwd, stream))) return rc;
    }
    return 0;
}

uint32_t nb200_train_plan_bytes(void) { return (uint32_t)sizeof(nb200_train_plan); }

}  // extern "C"
