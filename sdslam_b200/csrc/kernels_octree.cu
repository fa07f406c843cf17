// kernels_octree.cu -- keypoint culling of the ORB-SLAM2-style mode for sm_100a.
//
// NOT part of /root/reference: this is the public ORB-SLAM2 algorithm (raulmur/ORB_SLAM2,
// src/ORBextractor.cc: ComputeKeyPointsOctTree, DistributeOctTree, ExtractorNode::DivideNode), restated in
// oracle/sdorb_oracle.cc (compute_keypoints_octree, distribute_oct_tree) and a second time in tests/cv2_pipeline.py.
//
// ComputeKeyPointsOctTree, per level: cv::FAST on cells of about 30 pixels with iniThFAST, with minThFAST where that
// finds nothing; all keypoints of the level, cell by cell, go to DistributeOctTree, which splits the level into
// quadtree nodes until it has N of them and keeps the best keypoint of each.  The FAST kernel has already produced the
// keypoint map for min(iniThFAST, minThFAST) (a keypoint for iniThFAST is a keypoint for minThFAST whose score reaches
// iniThFAST) and gather_cells_kernel the ordered list of every cell with the cell's threshold choice applied.
//
// DistributeOctTree is std::list surgery as written, but every node enters the list at the front, so the list is always
// in descending creation order, a node's keypoints are a subset of the input in input order, and one "round" of the
// algorithm divides a set of nodes that is known up front: all nodes with more than one keypoint, in list order, or --
// once the next round could pass N -- in descending (size, address) order until N nodes exist.  So the kernel never
// moves keypoints: every keypoint carries the list position of its node, one CTA per (level, frame) runs the rounds as
// sweeps over the level's keypoints (child counts by shared-memory atomics on integers: order-independent, the result
// is bit-identical on every run), and only the node table (at most N + 3 entries) is rebuilt per round.
// Nodes of equal size are ordered by std::list node ADDRESSES in ORB-SLAM2, which is not a function of the input; like
// the oracle this kernel takes the later-created node first (addresses growing in creation order).
#include "kernels.cuh"

namespace sdorb {

constexpr int OCT_THREADS = 256;
constexpr int OCT_BORDER = SDORB_EDGE - 3;  // minBorderX = minBorderY = EDGE_THRESHOLD - 3

struct OctNode {
  int16_t x0, x1, y0, y1;  // UL = (x0, y0), BR = (x1, y1), relative to (minBorderX, minBorderY)
  int32_t size;            // vKeys.size()
};

__device__ __forceinline__ int oct_child(uint32_t e, const OctNode& nd) {
  const int x = SDORB_ENTRY_X(e) - OCT_BORDER, y = SDORB_ENTRY_Y(e) - OCT_BORDER;
  const int sx = nd.x0 + ((nd.x1 - nd.x0 + 1) >> 1), sy = nd.y0 + ((nd.y1 - nd.y0 + 1) >> 1);  // UL + ceil(extent / 2)
  return (x < sx ? 0 : 1) + (y < sy ? 0 : 2);  // n1, n2, n3, n4
}

__global__ void __launch_bounds__(OCT_THREADS) octree_kernel(const FrameGeom* __restrict__ geom, SelectBuffers buf, int max_cells,
                                                             int max_nodes) {
  extern __shared__ __align__(16) uint32_t smem[];
  pdl_enter();
  int* offs = reinterpret_cast<int*>(smem);                              // [max_cells + 1]
  OctNode* const nodes_a = reinterpret_cast<OctNode*>(offs + ((max_cells + 2) & ~1));  // [max_nodes] x 2
  OctNode* const nodes_b = nodes_a + max_nodes;
  int* ccnt = reinterpret_cast<int*>(nodes_b + max_nodes);               // [max_nodes][4] child sizes of the round
  int* ccnt_b = ccnt + 4 * max_nodes;                                    // [max_nodes][4] child sizes of the next table (filled on the way)
  int* cpos = ccnt_b + 4 * max_nodes;                                    // [max_nodes][4] new list positions of the children
  int* keep_pos = cpos + 4 * max_nodes;                                  // [max_nodes] new position of a node that stays; -1 = divided
  int* order = keep_pos + max_nodes;                                     // [max_nodes] nodes to divide, in processing order
  unsigned long long* best = reinterpret_cast<unsigned long long*>(order + max_nodes);  // [max_nodes]  (max_nodes is even)
  __shared__ int s_warp_sums[OCT_THREADS / 32];
  __shared__ int s_nalive, s_ncand, s_finish, s_sorted, s_total;

  const int level = blockIdx.y, frame = blockIdx.x;  // level-major grid: the longest CTAs first
  const LevelGeom& L = geom->lv[level];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_cells = (L.cols > 0 && L.rows > 0) ? L.cols * L.rows : 0;
  int32_t* out_count = buf.sel_count + (int64_t)frame * geom->nlevels + level;
  if (n_cells == 0 || L.list_cap_cell == 0) {
    if (tid == 0) *out_count = 0;
    return;
  }
  const int32_t* seen = buf.cell_seen + (int64_t)frame * geom->cells_total + L.cell_base;
  const uint32_t* lists = buf.cell_list + (int64_t)frame * geom->list_total + L.list_base;
  uint32_t* keys = buf.okeys + (int64_t)frame * geom->list_total + L.list_base;   // vToDistributeKeys
  uint16_t* knode = buf.onode + (int64_t)frame * geom->list_total + L.list_base;  // list position of each keypoint's node
  const int N = L.n_desired;

  // ---- vToDistributeKeys: the cells' lists concatenated row-major.  Exclusive scan of the cell counts:
  {
    const int per = (n_cells + OCT_THREADS - 1) / OCT_THREADS;
    const int c0 = tid * per, c1 = min(c0 + per, n_cells);
    int sum = 0;
    for (int c = c0; c < c1; ++c) sum += min(seen[c], L.list_cap_cell);
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp_sums[warp] = incl;
    __syncthreads();
    int base = incl - sum;
    for (int w2 = 0; w2 < warp; ++w2) base += s_warp_sums[w2];
    for (int c = c0; c < c1; ++c) {
      offs[c] = base;
      base += min(seen[c], L.list_cap_cell);
    }
    if (tid == OCT_THREADS - 1) {
      offs[n_cells] = base;
      s_total = base;
    }
    __syncthreads();
  }
  const int K = s_total;
  if (K == 0) {
    if (tid == 0) *out_count = 0;
    return;
  }
  // ---- initial nodes (DistributeOctTree's first loop) and the copy of the keypoints
  const int n_ini = L.n_ini;
  if (tid < n_ini) ccnt[tid] = 0;
  __syncthreads();
  for (int c = warp; c < n_cells; c += OCT_THREADS / 32) {
    const int o = offs[c], n = offs[c + 1] - o;
    const uint32_t* list = lists + (int64_t)c * L.list_cap_cell;
    for (int i = lane; i < n; i += 32) {
      const uint32_t e = list[i];
      keys[o + i] = e;
      const int ini = min((int)__fdiv_rn((float)(SDORB_ENTRY_X(e) - OCT_BORDER), L.h_x), n_ini - 1);  // vpIniNodes[kp.pt.x / hX]
      knode[o + i] = (uint16_t)ini;
      atomicAdd(&ccnt[ini], 1);
    }
  }
  __syncthreads();
  if (tid == 0) {
    int n = 0;
    for (int i = 0; i < n_ini; ++i) {
      keep_pos[i] = -1;
      if (ccnt[i] > 0) {  // empty initial nodes are erased
        OctNode nd;
        nd.x0 = (int16_t)(int)(L.h_x * (float)i);
        nd.x1 = (int16_t)(int)(L.h_x * (float)(i + 1));
        nd.y0 = 0;
        nd.y1 = (int16_t)(L.h - 2 * OCT_BORDER);  // maxY - minY
        nd.size = ccnt[i];
        keep_pos[i] = n;
        nodes_a[n++] = nd;
      }
    }
    s_nalive = n;
    s_sorted = 0;
    s_finish = 0;
  }
  __syncthreads();
  for (int k = tid; k < K; k += OCT_THREADS) knode[k] = (uint16_t)keep_pos[knode[k]];
  __syncthreads();

  // ---- the rounds.  ccnt[cb] holds the child sizes of the current table's nodes; the sweep that moves the keypoints to the
  // next table counts the children of that table into ccnt[cb ^ 1] on the way (one pass over the keypoints per round).
  uint32_t* ckey = reinterpret_cast<uint32_t*>(best);  // sort keys of the candidates of a sorted round (best is not used before the end)
  for (int i = tid; i < 4 * s_nalive; i += OCT_THREADS) ccnt[i] = 0;
  __syncthreads();
  for (int k = tid; k < K; k += OCT_THREADS) {
    const int i = knode[k];
    if (nodes_a[i].size > 1) atomicAdd(&ccnt[4 * i + oct_child(keys[k], nodes_a[i])], 1);
  }
  int cur = 0, cb = 0;
  while (true) {
    const OctNode* nd = cur ? nodes_b : nodes_a;
    OctNode* nn = cur ? nodes_a : nodes_b;
    const int* cc = cb ? ccnt_b : ccnt;
    int* cnext = cb ? ccnt : ccnt_b;
    const int n_alive = s_nalive;
    for (int i = tid; i < 4 * max_nodes; i += OCT_THREADS) cnext[i] = 0;
    // candidates in list order
    if (tid == 0) {
      int m = 0;
      for (int i = 0; i < n_alive; ++i)
        if (nd[i].size > 1) {
          ckey[m] = ((uint32_t)nd[i].size << 16) | (uint32_t)(0xFFFF - m);
          keep_pos[m++] = i;  // keep_pos doubles as the candidate list until the order is fixed
        }
      s_ncand = m;
    }
    __syncthreads();  // also: the child counts of this round are complete
    const int m = s_ncand;
    if (s_sorted) {
      // descending (size, address): rank every candidate; equal sizes -> the later-created node = the smaller position first
      for (int j = tid; j < m; j += OCT_THREADS) {
        const uint32_t kj = ckey[j];
        int r = 0;
        for (int q = 0; q < m; ++q) r += ckey[q] > kj ? 1 : 0;
        order[r] = keep_pos[j];
      }
    } else {
      for (int j = tid; j < m; j += OCT_THREADS) order[j] = keep_pos[j];
    }
    __syncthreads();
    if (tid == 0) {
      // which candidates are divided (all, or in sorted rounds until N nodes exist), in which order their children are created
      int size = n_alive, done = m, created = 0;
      for (int j = 0; j < m; ++j) {
        const int i = order[j];
        int nz = 0;
        for (int c = 0; c < 4; ++c) nz += cc[4 * i + c] > 0 ? 1 : 0;
        created += nz;
        size += nz - 1;
        if (s_sorted && size >= N) {
          done = j + 1;
          break;
        }
      }
      for (int i = 0; i < n_alive; ++i) keep_pos[i] = 0;
      for (int j = 0; j < done; ++j) keep_pos[order[j]] = -1;
      const int n_new = size;
      int overflow = n_new > max_nodes ? 1 : 0;
      int t = 0, n_expand = 0;
      for (int j = 0; j < done && !overflow; ++j) {
        const int i = order[j];
        const OctNode p = nd[i];
        const int16_t sx = (int16_t)(p.x0 + ((p.x1 - p.x0 + 1) >> 1)), sy = (int16_t)(p.y0 + ((p.y1 - p.y0 + 1) >> 1));
        for (int c = 0; c < 4; ++c) {
          const int cnt = cc[4 * i + c];
          if (cnt == 0) continue;
          const int pos = created - 1 - t;  // pushed to the front in creation order
          ++t;
          cpos[4 * i + c] = pos;
          OctNode ch;
          ch.x0 = (c & 1) ? sx : p.x0;
          ch.x1 = (c & 1) ? p.x1 : sx;
          ch.y0 = (c & 2) ? sy : p.y0;
          ch.y1 = (c & 2) ? p.y1 : sy;
          ch.size = cnt;
          nn[pos] = ch;
          n_expand += cnt > 1 ? 1 : 0;
        }
      }
      int kpos = created;
      for (int i = 0; i < n_alive && !overflow; ++i)
        if (keep_pos[i] == 0) {
          keep_pos[i] = kpos;
          nn[kpos++] = nd[i];
        }
      if (overflow) {
        atomicExch(buf.error_flag, 6);
        s_finish = 2;
      } else {
        s_nalive = n_new;
        if (n_new >= N || n_new == n_alive) s_finish = 1;
        else if (!s_sorted && n_new + 3 * n_expand > N) s_sorted = 1;
      }
    }
    __syncthreads();
    if (s_finish == 2) break;  // keeps the previous table
    const bool more = s_finish == 0;
    for (int k = tid; k < K; k += OCT_THREADS) {
      const int i = knode[k];
      const uint32_t e = keys[k];
      const int kp = keep_pos[i];
      const int np = kp >= 0 ? kp : cpos[4 * i + oct_child(e, nd[i])];
      knode[k] = (uint16_t)np;
      if (more && nn[np].size > 1) atomicAdd(&cnext[4 * np + oct_child(e, nn[np])], 1);
    }
    cur ^= 1;
    cb ^= 1;
    __syncthreads();
    if (s_finish) break;
  }

  // ---- the best keypoint of every node: the largest response, the first of them in vKeys order
  const int n_out = min(s_nalive, max_nodes);
  for (int i = tid; i < n_out; i += OCT_THREADS) best[i] = 0ull;
  __syncthreads();
  for (int k = tid; k < K; k += OCT_THREADS) {
    const unsigned long long v = ((unsigned long long)(keys[k] & 0xFFu) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)k);
    atomicMax(&best[knode[k]], v);
  }
  __syncthreads();
  uint32_t* sel = buf.sel + (int64_t)frame * geom->sel_total + L.sel_base;
  for (int i = tid; i < n_out; i += OCT_THREADS) sel[i] = keys[0xFFFFFFFFu - (uint32_t)(best[i] & 0xFFFFFFFFull)];
  if (tid == 0) *out_count = n_out;
}

static void octree_caps(const FrameGeom& g, int* max_cells, int* max_nodes) {
  int mc = 1, mn = 8;
  for (int l = 0; l < g.nlevels; ++l) {
    const LevelGeom& L = g.lv[l];
    mc = std::max(mc, L.cols * L.rows);
    const int slots = (l + 1 < g.nlevels ? g.lv[l + 1].sel_base : g.sel_total) - L.sel_base;
    mn = std::max(mn, slots);
  }
  *max_cells = mc;
  *max_nodes = (mn + 1) & ~1;  // even: keeps the 8-byte alignment of the shared-memory layout
}

static size_t octree_smem_bytes(int mc, int mn) {
  return sizeof(int) * (size_t)((mc + 2) & ~1) + sizeof(OctNode) * 2 * (size_t)mn + sizeof(int) * (size_t)(4 + 4 + 4 + 1 + 1) * mn +
         sizeof(unsigned long long) * (size_t)mn + 16;
}

void launch_octree(const FrameGeom* d_geom, const FrameGeom& g, const SelectBuffers& b, int nframes, cudaStream_t s) {
  int mc, mn;
  octree_caps(g, &mc, &mn);
  launch_pdl(octree_kernel, dim3(nframes, g.nlevels), dim3(OCT_THREADS), octree_smem_bytes(mc, mn), s, d_geom, b, mc, mn);
}

int configure_octree_kernel() {
  return (int)cudaFuncSetAttribute(octree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}

}  // namespace sdorb
