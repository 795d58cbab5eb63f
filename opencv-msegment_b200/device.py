"""Device-resident and asynchronous entry points of the C ABI (raw device / pinned pointers as ints).

Used by bench.py and by callers that keep frames in HBM (e.g. behind torch tensors: pass ``t.data_ptr()`` and
hand the torch stream to ``Context.set_stream``).  Thin ctypes forwarding only.
"""
import ctypes as C

from . import _lib as L


def params(sp=10.0, sr=10.0, max_level=1, termcrit=(3, 5, 1.0), lo_diff=2, min_size=0, color_dist=0, render_depth=0,
           connectivity=4, labels_type=0):
    return L.SegmentParams(float(sp), float(sr), int(max_level), int(termcrit[0]), int(termcrit[1]), float(termcrit[2]),
                           int(lo_diff), int(min_size), int(color_dist), int(render_depth), int(connectivity), int(labels_type))


def _p(v):
    return C.c_void_p(int(v)) if v else None


def synth(ctx, d_dst, step, w, h, seed):
    ctx.check(ctx._lib.msg_synth_bgr_dev(ctx._h, _p(d_dst), step, w, h, int(seed)))


def synth_rows(ctx, d_dst, step, w, full_h, row0, rows, seed):
    ctx.check(ctx._lib.msg_synth_bgr_rows_dev(ctx._h, _p(d_dst), step, w, full_h, row0, rows, int(seed)))


def segment(ctx, d_src, sstep, w, h, prm, d_filtered=0, fstep=0, d_labels=0, lstep=0, d_rendered=0, rstep=0, d_n=0):
    ctx.check(ctx._lib.msg_segment_dev(ctx._h, _p(d_src), sstep, w, h, C.byref(prm), _p(d_filtered), fstep, _p(d_labels),
                                       lstep, _p(d_rendered), rstep, _p(d_n)))


def meanshift(ctx, d_src, sstep, d_dst, dstep, w, h, sp, sr, max_level=1, termcrit=(3, 5, 1.0)):
    ctx.check(ctx._lib.msg_meanshift_filter_dev(ctx._h, _p(d_src), sstep, _p(d_dst), dstep, w, h, float(sp), float(sr),
                                                int(max_level), int(termcrit[0]), int(termcrit[1]), float(termcrit[2])))


def meanshift_strip(ctx, d_src_rows, sstep, halo_row0, halo_row1, d_dst, dstep, w, full_h, row0, row1, sp, sr,
                    max_level=1, termcrit=(3, 5, 1.0)):
    ctx.check(ctx._lib.msg_meanshift_filter_strip_dev(ctx._h, _p(d_src_rows), sstep, halo_row0, halo_row1, _p(d_dst), dstep,
                                                      w, full_h, row0, row1, float(sp), float(sr), int(max_level),
                                                      int(termcrit[0]), int(termcrit[1]), float(termcrit[2])))


def halo_rows(sp, max_level=1, termcrit=(3, 5, 1.0)):
    return L.load().msg_meanshift_halo_rows(float(sp), int(max_level), int(termcrit[0]), int(termcrit[1]))


def label_regions(ctx, d_bgr, step, d_labels, lstep, w, h, lo_diff, d_n=0, connectivity=4):
    ctx.check(ctx._lib.msg_label_regions_dev(ctx._h, _p(d_bgr), step, _p(d_labels), lstep, w, h, int(lo_diff),
                                             int(connectivity), _p(d_n)))


def label_strip(ctx, d_bgr_rows, step, d_labels, lstep, w, rows, row0, full_w, lo_diff):
    ctx.check(ctx._lib.msg_label_strip_dev(ctx._h, _p(d_bgr_rows), step, _p(d_labels), lstep, w, rows, row0, full_w,
                                           int(lo_diff)))


def seam_pairs(ctx, d_up_bgr, d_up_lab, d_lo_bgr, d_lo_lab, w, lo_diff, d_pairs, d_count):
    ctx.check(ctx._lib.msg_seam_pairs_dev(ctx._h, _p(d_up_bgr), _p(d_up_lab), _p(d_lo_bgr), _p(d_lo_lab), w, int(lo_diff),
                                          _p(d_pairs), _p(d_count)))


def apply_label_map(ctx, d_labels, lstep, w, rows, d_from, d_to, n):
    ctx.check(ctx._lib.msg_apply_label_map_dev(ctx._h, _p(d_labels), lstep, w, rows, _p(d_from), _p(d_to), int(n)))


def strip_rank(ctx, d_labels, lstep, w, rows, row0, full_w, d_count):
    ctx.check(ctx._lib.msg_strip_rank_dev(ctx._h, _p(d_labels), lstep, w, rows, row0, full_w, _p(d_count)))


def strip_query_dense(ctx, d_query, nq, w, rows, row0, full_w, offset, d_out):
    ctx.check(ctx._lib.msg_strip_query_dense_dev(ctx._h, _p(d_query), int(nq), w, rows, row0, full_w, int(offset), _p(d_out)))


def strip_apply_dense(ctx, d_labels, lstep, w, rows, row0, full_w, offset, d_rlab, d_rdense, nr):
    ctx.check(ctx._lib.msg_strip_apply_dense_dev(ctx._h, _p(d_labels), lstep, w, rows, row0, full_w, int(offset), _p(d_rlab),
                                                 _p(d_rdense), int(nr)))


def seam_quads(ctx, d_up_bgr, d_up_lab, d_up_rank1, d_lo_bgr, d_lo_lab, w, lo_diff, rows, row0, full_w, d_quads, d_count):
    ctx.check(ctx._lib.msg_seam_quads_dev(ctx._h, _p(d_up_bgr), _p(d_up_lab), _p(d_up_rank1), _p(d_lo_bgr), _p(d_lo_lab), w,
                                          int(lo_diff), rows, row0, full_w, _p(d_quads), _p(d_count)))


def strip_finalize_dense(ctx, d_labels, lstep, w, rows, row0, full_w, offset, d_frm, d_dense, n_map, frm_lo):
    ctx.check(ctx._lib.msg_strip_finalize_dense_dev(ctx._h, _p(d_labels), lstep, w, rows, row0, full_w, int(offset), _p(d_frm),
                                                    _p(d_dense), int(n_map), int(frm_lo)))


def shard_plan(w, h, n_strips, sp, max_level=1, termcrit=(3, 5, 1.0)):
    """msg_shard_plan_make -> (halo_rows, [(row0, row1)], [(halo0, halo1)])"""
    p = L.ShardPlan()
    rc = L.load().msg_shard_plan_make(int(w), int(h), int(n_strips), float(sp), int(max_level), int(termcrit[0]), int(termcrit[1]),
                                      C.byref(p))
    if rc != L.MSG_OK:
        raise ValueError("msg_shard_plan_make(%d x %d, %d strips, max_level %d) failed: %d" % (w, h, n_strips, max_level, rc))
    return (p.halo_rows, [(p.row0[k], p.row1[k]) for k in range(n_strips)], [(p.halo0[k], p.halo1[k]) for k in range(n_strips)])


def strip_resolve_dense(ctx, d_gathered, n_strips, w, row0s, d_tables, tables_ints):
    arr = (C.c_int * n_strips)(*[int(r) for r in row0s])
    ctx.check(ctx._lib.msg_strip_resolve_dense_dev(ctx._h, _p(d_gathered), int(n_strips), int(w), arr, _p(d_tables), int(tables_ints)))


def strip_finalize_tables(ctx, d_labels, lstep, w, rows, row0, full_w, strip, n_strips, d_tables):
    ctx.check(ctx._lib.msg_strip_finalize_tables_dev(ctx._h, _p(d_labels), lstep, w, rows, row0, full_w, int(strip), int(n_strips),
                                                     _p(d_tables)))


def strip_merge_stats(ctx, d_bgr, step, d_labels, lstep, w, rows, d_up_row_labels, n_total, d_area, d_sum, d_pairs, pair_cap, d_npairs):
    ctx.check(ctx._lib.msg_strip_merge_stats_dev(ctx._h, _p(d_bgr), step, _p(d_labels), lstep, w, rows, _p(d_up_row_labels), int(n_total),
                                                 _p(d_area), _p(d_sum), _p(d_pairs), int(pair_cap), _p(d_npairs)))


def strip_merge_finish(ctx, d_labels, lstep, w, rows, full_pixels, n_total, d_area, d_sum, d_all_pairs, n_all_pairs, min_size, color_dist,
                       d_n_out=0):
    ctx.check(ctx._lib.msg_strip_merge_finish_dev(ctx._h, _p(d_labels), lstep, w, rows, int(full_pixels), int(n_total), _p(d_area),
                                                  _p(d_sum), _p(d_all_pairs), int(n_all_pairs), int(min_size), int(color_dist), _p(d_n_out)))


def connected_components(ctx, d_mask, step, d_labels, lstep, w, h, connectivity=8, d_n=0):
    ctx.check(ctx._lib.msg_connected_components_dev(ctx._h, _p(d_mask), step, _p(d_labels), lstep, w, h, int(connectivity),
                                                    _p(d_n)))


def merge_regions(ctx, d_bgr, step, d_labels, lstep, w, h, min_size, color_dist, d_n=0):
    ctx.check(ctx._lib.msg_merge_regions_dev(ctx._h, _p(d_bgr), step, _p(d_labels), lstep, w, h, int(min_size),
                                             int(color_dist), _p(d_n)))


def render_labels(ctx, d_labels, lstep, d_dst, dstep, w, h, depth, d_colors=0):
    ctx.check(ctx._lib.msg_render_labels_dev(ctx._h, _p(d_labels), lstep, _p(d_dst), dstep, w, h, int(depth), _p(d_colors)))


def watershed_batch(ctx, d_bgr, step, image_stride, d_markers, mstep, markers_stride, w, h, count, d_pops=0):
    ctx.check(ctx._lib.msg_watershed_batch_dev(ctx._h, _p(d_bgr), step, image_stride, _p(d_markers), mstep, markers_stride, w, h,
                                               int(count), _p(d_pops)))


# ---- asynchronous host-buffer interface (pinned memory)

def alloc_pinned(nbytes):
    p = L.load().msg_alloc_pinned(int(nbytes))
    if not p:
        raise MemoryError("msg_alloc_pinned(%d) failed" % nbytes)
    return p


def free_pinned(p):
    L.load().msg_free_pinned(C.c_void_p(p))


def submit_segment(ctx, src, sstep, w, h, prm, filtered=0, fstep=0, labels=0, lstep=0, rendered=0, rstep=0):
    t = C.c_int()
    ctx.check(ctx._lib.msg_submit_segment(ctx._h, _p(src), sstep, w, h, C.byref(prm), _p(filtered), fstep, _p(labels), lstep,
                                          _p(rendered), rstep, C.byref(t)))
    return t.value


def wait(ctx, ticket):
    n = C.c_int32()
    ctx.check(ctx._lib.msg_wait(ctx._h, int(ticket), C.byref(n)))
    return n.value
