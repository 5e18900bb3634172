// Mock containers with the interface FEDDLib's BCBuilder (core/General/BCBuilder_def.hpp) and BlockMatrix::merge read: a CSR
// backed fill-complete Matrix with local row views, Map, Domain with boundary flags, (Block)MultiVector, BlockMatrix.
// Test infrastructure only (oracle/_ref): lets the reference's own routines compile unmodified without Trilinos.
#pragma once
#include <algorithm>
#include <stdexcept>
#include <string>
#include <vector>

#include "teuchos_mock.hpp"

namespace FEDD {

template <class LO, class GO, class NO>
class Map {
  public:
    Map(const GO *gids, std::size_t n) : gids_(gids, gids + n) {}
    GO getGlobalElement(LO i) const { return gids_.at(i); }
    std::size_t getNodeNumElements() const { return gids_.size(); }
  private:
    std::vector<GO> gids_;
};

template <class SC, class LO, class GO, class NO>
class MultiVector {
  public:
    typedef Map<LO, GO, NO> Map_Type;
    typedef Teuchos::RCP<const Map_Type> MapConstPtr_Type;
    MultiVector(SC *data, std::size_t n) : p_(data), n_(n) {}
    explicit MultiVector(MapConstPtr_Type) : p_(nullptr), n_(0) { throw std::logic_error("mock MultiVector(map): Neumann branch is outside the checker"); }
    unsigned getNumVectors() const { return 1; }
    Teuchos::ArrayRCP<SC> getDataNonConst(int) const { return Teuchos::ArrayRCP<SC>(p_, n_); }
    template <class V> void exportFromVector(const V &, bool, const std::string &) { throw std::logic_error("mock"); }
    void update(SC, const MultiVector &, SC) { throw std::logic_error("mock"); }
  private:
    SC *p_;
    std::size_t n_;
};

template <class SC, class LO, class GO, class NO>
class BlockMultiVector {
  public:
    typedef MultiVector<SC, LO, GO, NO> MultiVector_Type;
    typedef Teuchos::RCP<MultiVector_Type> MultiVectorPtr_Type;
    int size() const { return (int)blocks_.size(); }
    unsigned getNumVectors() const { return 1; }
    MultiVectorPtr_Type getBlock(int i) const { return blocks_.at(i); }
    MultiVectorPtr_Type getBlockNonConst(int i) const { return blocks_.at(i); }
    std::vector<MultiVectorPtr_Type> blocks_;
};

template <class SC, class LO, class GO, class NO>
class Domain {
  public:
    typedef Map<LO, GO, NO> Map_Type;
    typedef Teuchos::RCP<const Map_Type> MapConstPtr_Type;
    int getDimension() const { return dim_; }
    std::string getFEType() const { return "P2"; }
    Teuchos::RCP<std::vector<int> > getBCFlagUnique() const { return flags_; }
    Teuchos::RCP<std::vector<std::vector<double> > > getPointsUnique() const { return points_; }
    MapConstPtr_Type getMapUnique() const { return mapUnique_; }
    MapConstPtr_Type getMapRepeated() const { return mapUnique_; }
    MapConstPtr_Type getMapVecFieldUnique() const { return mapUnique_; }
    MapConstPtr_Type getMapVecFieldRepeated() const { return mapUnique_; }
    int dim_ = 0;
    Teuchos::RCP<std::vector<int> > flags_;
    Teuchos::RCP<std::vector<std::vector<double> > > points_;
    MapConstPtr_Type mapUnique_;
};

template <class SC, class LO, class GO, class NO>
class FE {
  public:
    template <class D> void addFE(const D &) {}
    template <class... A> void assemblySurfaceIntegralFlag(A &&...) { throw std::logic_error("mock FE: Neumann branch is outside the checker"); }
};

// fill-complete CSR matrix on one rank: local rows / local columns, row and column maps
template <class SC, class LO, class GO, class NO>
class Matrix {
  public:
    typedef Map<LO, GO, NO> Map_Type;
    typedef Teuchos::RCP<Map_Type> MapPtr_Type;
    typedef Teuchos::RCP<const Map_Type> MapConstPtr_Type;
    Matrix(const long long *rowptr, const LO *colind, SC *values, MapConstPtr_Type rowMap, MapConstPtr_Type colMap)
        : rp_(rowptr), ci_(colind), v_(values), row_(rowMap), col_(colMap), filling_(false), resumes_(0), completes_(0) {}
    void resumeFill() { filling_ = true; resumes_++; }
    void fillComplete(MapConstPtr_Type, MapConstPtr_Type) { filling_ = false; completes_++; }
    MapConstPtr_Type getMap(const std::string &which = "row") const { return which == "col" ? col_ : row_; }
    void getLocalRowView(LO r, Teuchos::ArrayView<const LO> &idx, Teuchos::ArrayView<const SC> &val) const
    {
        idx = Teuchos::ArrayView<const LO>(ci_ + rp_[r], (std::size_t)(rp_[r + 1] - rp_[r]));
        val = Teuchos::ArrayView<const SC>(v_ + rp_[r], (std::size_t)(rp_[r + 1] - rp_[r]));
    }
    template <class IV, class VV>
    void replaceLocalValues(LO r, const IV &idx, const VV &val)
    {
        if (!filling_) throw std::runtime_error("replaceLocalValues on a fill-complete matrix");
        for (std::size_t k = 0; k < idx.size(); k++) {
            const LO *b = ci_ + rp_[r], *e = ci_ + rp_[r + 1];
            const LO *it = std::find(b, e, idx[k]);
            if (it == e) throw std::runtime_error("replaceLocalValues: column not in the row");
            v_[rp_[r] + (it - b)] = val[k];
        }
    }
    const long long *rp_;
    const LO *ci_;
    SC *v_;
    MapConstPtr_Type row_, col_;
    bool filling_;
    int resumes_, completes_;
};

template <class SC, class LO, class GO, class NO>
class BlockMatrix {
  public:
    typedef Matrix<SC, LO, GO, NO> Matrix_Type;
    typedef Teuchos::RCP<Matrix_Type> MatrixPtr_Type;
    explicit BlockMatrix(int n) : n_(n), blocks_((std::size_t)n * n) {}
    int size() const { return n_; }
    bool blockExists(int i, int j) const { return !blocks_.at((std::size_t)i * n_ + j).is_null(); }
    MatrixPtr_Type getBlock(int i, int j) const { return blocks_.at((std::size_t)i * n_ + j); }
    void addBlock(const MatrixPtr_Type &m, int i, int j) { blocks_.at((std::size_t)i * n_ + j) = m; }
    int n_;
    std::vector<MatrixPtr_Type> blocks_;
};

} // namespace FEDD
