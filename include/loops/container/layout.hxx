/**
 * @file layout.hxx
 * @brief Tile/atom layout views -- the contract every schedule consumes.
 *
 * Keeps the reference's layout-view contract verbatim (reference
 * include/loops/container/layout.hxx:16-55):
 *
 *     num_tiles()  num_atoms()  tile_begin(t)  tile_end(t)  tile_size(t)
 *     tile_end_iter()            [+ tile_of(a) where it is cheap]
 *
 * with `tile_begin(0) == 0`, `tile_end(T-1) == num_atoms()` and `tile_end`
 * non-decreasing. Views are trivially copyable, non-owning and passed BY VALUE
 * into kernels. Same type names and constructor signatures as the reference:
 *
 *   layout::csr  (:87-149)   layout::bcsr (:239-285)  layout::csc (:312-359)
 *   layout::coo  (:385-421)  layout::ell  (:443-496)  layout::dia (:166-217)
 *   layout::flat_uniform_occupancy<K, base>  (container/partitioning.hxx:71-141)
 *
 * Implementation here is two generic views -- an offsets-array view and an
 * affine ("every tile has `pitch` atoms") view -- that the named layouts derive
 * from; `tile_end_iter()` for arithmetic layouts is a small affine iterator
 * (no thrust dependency). Every view can also describe itself to the C ABI via
 * `descriptor()` (include/loopsb.h: loopsb_layout_t) so the host wrappers in
 * loops/algorithms/spmv/ can hand the work to the sm_100a kernels.
 */
#pragma once

#include <cstddef>
#include <cstdint>

#include <loops/range.hxx>
#include <loopsb.h>

namespace loops {

/// Random-access "iterator" over an arithmetic sequence: it[k] = min(cap,
/// first + k * stride). Used as `tile_end_iter()` by layouts whose tile ends
/// are a formula rather than an array, and as the counting iterator of the
/// schedules (stride 1, no cap).
template <typename value_type>
struct affine_iterator {
  using value_t = value_type;
  using difference_t = long long;
  value_t first;
  value_t stride;
  value_t cap;  ///< upper clamp; ignored when `capped` is false
  bool capped;

  LOOPS_HD affine_iterator() : first(0), stride(1), cap(0), capped(false) {}
  LOOPS_HD explicit affine_iterator(value_t f, value_t s = value_t(1))
      : first(f), stride(s), cap(0), capped(false) {}
  LOOPS_HD affine_iterator(value_t f, value_t s, value_t c)
      : first(f), stride(s), cap(c), capped(true) {}

  LOOPS_HD value_t operator[](difference_t k) const {
    const value_t v = static_cast<value_t>(first + static_cast<value_t>(k) * stride);
    return (capped && cap < v) ? cap : v;
  }
  LOOPS_HD value_t operator*() const { return (*this)[0]; }
  LOOPS_HD affine_iterator operator+(difference_t k) const {
    affine_iterator r(*this);
    r.first = static_cast<value_t>(first + static_cast<value_t>(k) * stride);
    return r;
  }
  LOOPS_HD affine_iterator& operator++() {
    first = static_cast<value_t>(first + stride);
    return *this;
  }
};

/// `counting_iterator<T>(n)[k] == n + k`.
template <typename T>
using counting_iterator = affine_iterator<T>;

namespace layout {
namespace detail {

/// Tiles delimited by a monotone offsets array of length num_tiles + 1.
template <typename tile_id_type, typename atom_id_type, int kind_tag>
struct offsets_view {
  using tile_id_t = tile_id_type;
  using atom_id_t = atom_id_type;
  using tile_end_iterator_t = atom_id_t const*;

  atom_id_t const* offsets_;
  tile_id_t n_tiles_;
  atom_id_t n_atoms_;

  LOOPS_HD offsets_view() : offsets_(nullptr), n_tiles_(0), n_atoms_(0) {}
  LOOPS_HD offsets_view(atom_id_t const* offsets,
                        tile_id_t num_tiles,
                        atom_id_t num_atoms)
      : offsets_(offsets), n_tiles_(num_tiles), n_atoms_(num_atoms) {}

  LOOPS_HD tile_id_t num_tiles() const { return n_tiles_; }
  LOOPS_HD atom_id_t num_atoms() const { return n_atoms_; }
  LOOPS_HD atom_id_t tile_begin(tile_id_t t) const { return offsets_[t]; }
  LOOPS_HD atom_id_t tile_end(tile_id_t t) const { return offsets_[t + 1]; }
  LOOPS_HD atom_id_t tile_size(tile_id_t t) const {
    return offsets_[t + 1] - offsets_[t];
  }
  LOOPS_HD tile_end_iterator_t tile_end_iter() const { return offsets_ + 1; }

  /// Owner of atom `a`: the first tile whose end lies beyond `a` (empty tiles
  /// are skipped because their end does not exceed their begin).
  LOOPS_HD tile_id_t tile_of(atom_id_t a) const {
    tile_id_t first = 0;
    tile_id_t count = n_tiles_;
    while (count > 0) {
      const tile_id_t half = count / 2;
      if (offsets_[first + half + 1] <= a) {
        first += half + 1;
        count -= half + 1;
      } else {
        count = half;
      }
    }
    return first;
  }

  /// C-ABI description of this view (include/loopsb.h).
  loopsb_layout_t descriptor() const {
    static_assert(sizeof(atom_id_type) == 4,
                  "the C ABI carries int32 ids (the reference uses int)");
    loopsb_layout_t d{};
    d.kind = kind_tag;
    d.offsets = reinterpret_cast<const int32_t*>(offsets_);
    d.num_tiles = static_cast<int32_t>(n_tiles_);
    d.num_atoms = static_cast<int32_t>(n_atoms_);
    d.pitch = 0;
    return d;
  }
};

/// Every tile owns exactly `pitch` consecutive atoms.
template <typename tile_id_type, typename atom_id_type, int kind_tag>
struct pitch_view {
  using tile_id_t = tile_id_type;
  using atom_id_t = atom_id_type;
  using tile_end_iterator_t = affine_iterator<atom_id_t>;

  tile_id_t n_tiles_;
  atom_id_t pitch_;

  LOOPS_HD pitch_view() : n_tiles_(0), pitch_(0) {}
  LOOPS_HD pitch_view(tile_id_t num_tiles, atom_id_t pitch)
      : n_tiles_(num_tiles), pitch_(pitch) {}

  LOOPS_HD tile_id_t num_tiles() const { return n_tiles_; }
  LOOPS_HD atom_id_t num_atoms() const {
    return static_cast<atom_id_t>(n_tiles_) * pitch_;
  }
  LOOPS_HD atom_id_t tile_begin(tile_id_t t) const {
    return static_cast<atom_id_t>(t) * pitch_;
  }
  LOOPS_HD atom_id_t tile_end(tile_id_t t) const {
    return static_cast<atom_id_t>(t + 1) * pitch_;
  }
  LOOPS_HD atom_id_t tile_size(tile_id_t) const { return pitch_; }
  LOOPS_HD tile_end_iterator_t tile_end_iter() const {
    return tile_end_iterator_t(pitch_, pitch_);  // it[k] = (k + 1) * pitch
  }
  LOOPS_HD tile_id_t tile_of(atom_id_t a) const {
    return static_cast<tile_id_t>(a / pitch_);
  }

  loopsb_layout_t descriptor() const {
    loopsb_layout_t d{};
    d.kind = kind_tag;
    d.offsets = nullptr;
    d.num_tiles = static_cast<int32_t>(n_tiles_);
    d.num_atoms = static_cast<int32_t>(num_atoms());
    d.pitch = static_cast<int32_t>(pitch_);
    return d;
  }
};

}  // namespace detail

/// CSR: tile = row, atom = stored nonzero.
template <typename tile_id_type, typename atom_id_type>
struct csr : detail::offsets_view<tile_id_type, atom_id_type, LOOPSB_LAYOUT_CSR> {
  using base_t = detail::offsets_view<tile_id_type, atom_id_type, LOOPSB_LAYOUT_CSR>;
  using base_t::base_t;
  LOOPS_HD csr() : base_t() {}
};

/// CSC: tile = column, atom = stored nonzero.
template <typename tile_id_type, typename atom_id_type>
struct csc : detail::offsets_view<tile_id_type, atom_id_type, LOOPSB_LAYOUT_CSC> {
  using base_t = detail::offsets_view<tile_id_type, atom_id_type, LOOPSB_LAYOUT_CSC>;
  using base_t::base_t;
  LOOPS_HD csc() : base_t() {}
};

/// BCSR: tile = block-row, atom = stored R x C block.
template <typename tile_id_type, typename atom_id_type>
struct bcsr : detail::offsets_view<tile_id_type, atom_id_type, LOOPSB_LAYOUT_BCSR> {
  using base_t = detail::offsets_view<tile_id_type, atom_id_type, LOOPSB_LAYOUT_BCSR>;
  using base_t::base_t;
  LOOPS_HD bcsr() : base_t() {}
};

/// ELL: tile = row, `pitch` slots per row (padding slots carry column -1).
template <typename tile_id_type, typename atom_id_type>
struct ell : detail::pitch_view<tile_id_type, atom_id_type, LOOPSB_LAYOUT_ELL> {
  using base_t = detail::pitch_view<tile_id_type, atom_id_type, LOOPSB_LAYOUT_ELL>;
  using base_t::base_t;
  LOOPS_HD ell() : base_t() {}
};

/// DIA: tile = row, one slot per stored diagonal.
template <typename tile_id_type, typename atom_id_type>
struct dia : detail::pitch_view<tile_id_type, atom_id_type, LOOPSB_LAYOUT_DIA> {
  using base_t = detail::pitch_view<tile_id_type, atom_id_type, LOOPSB_LAYOUT_DIA>;
  using base_t::base_t;
  LOOPS_HD dia() : base_t() {}
};

/// COO: every nonzero is its own tile (tiles == atoms == nnz).
template <typename tile_id_type, typename atom_id_type>
struct coo {
  using tile_id_t = tile_id_type;
  using atom_id_t = atom_id_type;
  using tile_end_iterator_t = counting_iterator<atom_id_t>;

  atom_id_t n_nzs_;

  LOOPS_HD coo() : n_nzs_(0) {}
  LOOPS_HD explicit coo(atom_id_t nnz) : n_nzs_(nnz) {}

  LOOPS_HD tile_id_t num_tiles() const { return static_cast<tile_id_t>(n_nzs_); }
  LOOPS_HD atom_id_t num_atoms() const { return n_nzs_; }
  LOOPS_HD atom_id_t tile_begin(tile_id_t t) const {
    return static_cast<atom_id_t>(t);
  }
  LOOPS_HD atom_id_t tile_end(tile_id_t t) const {
    return static_cast<atom_id_t>(t) + 1;
  }
  LOOPS_HD atom_id_t tile_size(tile_id_t) const { return atom_id_t(1); }
  LOOPS_HD tile_end_iterator_t tile_end_iter() const {
    return tile_end_iterator_t(atom_id_t(1));  // it[k] = k + 1
  }
  LOOPS_HD tile_id_t tile_of(atom_id_t a) const {
    return static_cast<tile_id_t>(a);
  }

  loopsb_layout_t descriptor() const {
    loopsb_layout_t d{};
    d.kind = LOOPSB_LAYOUT_COO;
    d.offsets = nullptr;
    d.num_tiles = static_cast<int32_t>(n_nzs_);
    d.num_atoms = static_cast<int32_t>(n_nzs_);
    d.pitch = 1;
    return d;
  }
};

/// Partitioner adaptor: re-tiles any base layout into windows of K atoms
/// (last window short). `base()` keeps the original view reachable so a kernel
/// can recover the original tile of an atom with `base().tile_of(a)`.
template <std::size_t K, typename base_layout_type>
struct flat_uniform_occupancy {
  static_assert(K > 0, "flat_uniform_occupancy: K must be positive.");
  using base_layout_t = base_layout_type;
  using tile_id_t = typename base_layout_t::tile_id_t;
  using atom_id_t = typename base_layout_t::atom_id_t;
  using tile_end_iterator_t = affine_iterator<atom_id_t>;

  static constexpr atom_id_t kAtomsPerTile = static_cast<atom_id_t>(K);

  base_layout_t base_;

  LOOPS_HD flat_uniform_occupancy() : base_() {}
  LOOPS_HD explicit flat_uniform_occupancy(base_layout_t base) : base_(base) {}

  LOOPS_HD const base_layout_t& base() const { return base_; }

  LOOPS_HD tile_id_t num_tiles() const {
    return static_cast<tile_id_t>((base_.num_atoms() + kAtomsPerTile - 1) /
                                  kAtomsPerTile);
  }
  LOOPS_HD atom_id_t num_atoms() const { return base_.num_atoms(); }
  LOOPS_HD atom_id_t tile_begin(tile_id_t t) const {
    return static_cast<atom_id_t>(t) * kAtomsPerTile;
  }
  LOOPS_HD atom_id_t tile_end(tile_id_t t) const {
    const atom_id_t e = static_cast<atom_id_t>(t + 1) * kAtomsPerTile;
    const atom_id_t total = base_.num_atoms();
    return e < total ? e : total;
  }
  LOOPS_HD atom_id_t tile_size(tile_id_t t) const {
    return tile_end(t) - tile_begin(t);
  }
  LOOPS_HD tile_end_iterator_t tile_end_iter() const {
    return tile_end_iterator_t(kAtomsPerTile, kAtomsPerTile, base_.num_atoms());
  }
  LOOPS_HD tile_id_t tile_of(atom_id_t a) const {
    return static_cast<tile_id_t>(a / kAtomsPerTile);
  }

  loopsb_layout_t descriptor() const {
    loopsb_layout_t d{};
    d.kind = LOOPSB_LAYOUT_FLAT;
    d.offsets = base_.descriptor().offsets;   // the base CSR row offsets (tile_of of the base layout)
    d.num_tiles = static_cast<int32_t>(num_tiles());
    d.num_atoms = static_cast<int32_t>(num_atoms());
    d.pitch = static_cast<int32_t>(K);
    return d;
  }
};

}  // namespace layout
}  // namespace loops
