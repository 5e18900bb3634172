// Mock containers with the interface FEDDLib's BlockMatrix::merge / BlockMap::merge read (core/LinearAlgebra/BlockMatrix_def.hpp
// :119-270, BlockMap_def.hpp:41-92): Map with element list and index extrema, a Matrix that is either a fill-complete local CSR
// (the blocks) or an insert-accumulating global matrix (the merged one).  Test infrastructure only (oracle/_ref).
#pragma once
#include <algorithm>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "teuchos_mock.hpp"

namespace Teuchos {
enum EVerbosityLevel { VERB_DEFAULT, VERB_EXTREME };
template <class T, int N>
struct Tuple {
    T v[N];
    T &operator[](int i) { return v[i]; }
    const T &operator[](int i) const { return v[i]; }
};
template <class T> Tuple<T, 2> tuple(const T &a, const T &b) { Tuple<T, 2> t; t.v[0] = a; t.v[1] = b; return t; }
struct MockComm { };
} // namespace Teuchos

namespace FEDD {

template <class LO, class GO, class NO>
class Map {
  public:
    typedef Map<LO, GO, NO> Map_Type;
    typedef Teuchos::RCP<Map_Type> MapPtr_Type;
    typedef Teuchos::RCP<const Map_Type> MapConstPtr_Type;
    typedef Teuchos::MockComm Comm_Type;
    typedef Teuchos::RCP<Comm_Type> CommPtr_Type;
    typedef Teuchos::RCP<const Comm_Type> CommConstPtr_Type;
    Map(const GO *gids, std::size_t n) : gids_(gids, gids + n) {}
    Map(const std::string &, GO, const Teuchos::ArrayView<GO> &list, GO, CommConstPtr_Type) : gids_(list.getRawPtr(), list.getRawPtr() + list.size()) {}
    Map(const std::string &, GO, const Teuchos::ArrayView<const GO> &list, GO, CommConstPtr_Type) : gids_(list.getRawPtr(), list.getRawPtr() + list.size()) {}
    GO getGlobalElement(LO i) const { return gids_.at(i); }
    std::size_t getNodeNumElements() const { return gids_.size(); }
    Teuchos::ArrayView<const GO> getNodeElementList() const { return Teuchos::ArrayView<const GO>(gids_.data(), gids_.size()); }
    GO getMaxAllGlobalIndex() const { return gids_.empty() ? GO(-1) : *std::max_element(gids_.begin(), gids_.end()); }
    LO getMaxLocalIndex() const { return (LO)gids_.size() - 1; }
    CommConstPtr_Type getComm() const { return CommConstPtr_Type(); }
    CommPtr_Type getCommNonConst() { return CommPtr_Type(); }
    std::string getUnderlyingLib() const { return "Tpetra"; }
    std::vector<GO> gids_;
};

template <class SC, class LO, class GO, class NO>
class MultiVector { };
template <class SC, class LO, class GO, class NO>
class BlockMultiVector { };

template <class SC, class LO, class GO, class NO>
class Matrix {
  public:
    typedef Map<LO, GO, NO> Map_Type;
    typedef Teuchos::RCP<Map_Type> MapPtr_Type;
    typedef Teuchos::RCP<const Map_Type> MapConstPtr_Type;
    typedef MultiVector<SC, LO, GO, NO> MultiVector_Type;
    typedef Teuchos::RCP<MultiVector_Type> MultiVectorPtr_Type;
    typedef Teuchos::RCP<const MultiVector_Type> MultiVectorConstPtr_Type;
    // a fill-complete block: local CSR with row and column maps
    Matrix(const long long *rowptr, const LO *colind, const SC *values, MapConstPtr_Type rowMap, MapConstPtr_Type colMap)
        : rp_(rowptr), ci_(colind), v_(values), row_(rowMap), col_(colMap), local_(true) {}
    // Matrix(map, maxNumEntries): globally indexed, fill-active (Matrix_def.hpp:46-51)
    Matrix(MapConstPtr_Type rowMap, LO) : rp_(nullptr), ci_(nullptr), v_(nullptr), row_(rowMap), local_(false), ins_(rowMap->getNodeNumElements())
    {
        for (std::size_t k = 0; k < rowMap->gids_.size(); k++) lrow_[rowMap->gids_[k]] = k;
    }
    explicit Matrix(const Teuchos::RCP<Matrix> &) { throw std::logic_error("mock Matrix copy"); }
    MapConstPtr_Type getMap(const std::string &which = "row") const { return which == "col" ? col_ : row_; }
    bool isLocallyIndexed() const { return local_; }
    std::size_t getNodeNumRows() const { return row_->getNodeNumElements(); }
    LO getGlobalMaxNumRowEntries() const
    {
        LO m = 0;
        for (std::size_t r = 0; r < getNodeNumRows(); r++) m = std::max<LO>(m, (LO)(rp_[r + 1] - rp_[r]));
        return m;
    }
    void getLocalRowView(LO r, Teuchos::ArrayView<const LO> &idx, Teuchos::ArrayView<const SC> &val) const
    {
        idx = Teuchos::ArrayView<const LO>(ci_ + rp_[r], (std::size_t)(rp_[r + 1] - rp_[r]));
        val = Teuchos::ArrayView<const SC>(v_ + rp_[r], (std::size_t)(rp_[r + 1] - rp_[r]));
    }
    template <class IV, class VV>
    void insertGlobalValues(GO row, const IV &cols, const VV &vals)
    {
        auto it = lrow_.find(row);
        if (it == lrow_.end()) throw std::runtime_error("insertGlobalValues: row not in the row map");
        for (std::size_t k = 0; k < cols.size(); k++) ins_[it->second].push_back(std::make_pair((GO)cols[k], (SC)vals[k]));
    }
    // fillComplete: sort the entries of every row by column gid (stable) and sum duplicates in insertion order
    void fillComplete(MapConstPtr_Type, MapConstPtr_Type)
    {
        for (auto &r : ins_) {
            std::stable_sort(r.begin(), r.end(), [](const std::pair<GO, SC> &a, const std::pair<GO, SC> &b) { return a.first < b.first; });
            std::vector<std::pair<GO, SC> > m;
            for (const auto &e : r) {
                if (!m.empty() && m.back().first == e.first) m.back().second += e.second;
                else m.push_back(e);
            }
            r.swap(m);
        }
    }
    const long long *rp_;
    const LO *ci_;
    const SC *v_;
    MapConstPtr_Type row_, col_;
    bool local_;
    std::vector<std::vector<std::pair<GO, SC> > > ins_;
    std::map<GO, std::size_t> lrow_;
};

} // namespace FEDD
