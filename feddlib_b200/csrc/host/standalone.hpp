// standalone.hpp -- the handful of FEDDLib / Teuchos types FE_b200 reads, for programs that use the engine WITHOUT a
// Trilinos build (the C++ end-to-end benchmark bench_fe_b200.cpp, examples, smoke tests of the host layer).  With
// FEDDLib present, include its headers instead (and define FEDD_B200_TRILINOS): FE_b200.hpp is written against exactly
// this surface -- Domain::getDimension/getFEType/getElementsC/getPointsRepeated/getMapRepeated (Domain_decl.hpp:21-247),
// Elements::numberElements/getElement (Elements.hpp:21-109), FiniteElement::getVectorNodeList (FiniteElement.hpp:17-110),
// Map::getGlobalElement/getNodeNumElements (Map_decl.hpp:27-109), MultiVector::getData (MultiVector_decl.hpp).
// The Matrix here only keeps what the seat step hands over: the shared CSR structure and the pooled value buffer.
#pragma once
#include <cstddef>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace Teuchos {
template <class T>
class RCP {
  public:
    RCP() {}
    explicit RCP(T *p) : p_(p) {}
    template <class U> RCP(const RCP<U> &o) : p_(o.shared()) {}
    T *operator->() const { return p_.get(); }
    T &operator*() const { return *p_; }
    T *get() const { return p_.get(); }
    bool is_null() const { return !p_; }
    const std::shared_ptr<T> &shared() const { return p_; }
  private:
    std::shared_ptr<T> p_;
};
template <class T> RCP<T> rcp(T *p) { return RCP<T>(p); }

template <class T>
class ArrayRCP {   // non-owning view, enough for MultiVector::getData
  public:
    ArrayRCP() : p_(nullptr), n_(0) {}
    ArrayRCP(T *p, std::size_t n) : p_(p), n_(n) {}
    std::size_t size() const { return n_; }
    T &operator[](std::size_t i) const { return p_[i]; }
  private:
    T *p_;
    std::size_t n_;
};
} // namespace Teuchos

namespace FEDD {
namespace b200 {
template <class SC, class LO, class GO> struct LocalCsr;
}

template <class LO, class GO, class NO>
class Map {
  public:
    Map(const GO *gids, std::size_t n) : gids_(gids, gids + n) {}
    GO getGlobalElement(LO i) const { return gids_[(std::size_t)i]; }
    std::size_t getNodeNumElements() const { return gids_.size(); }
  private:
    std::vector<GO> gids_;
};

class FiniteElement {
  public:
    FiniteElement() {}
    FiniteElement(const int *nodes, int n) : nodes_(nodes, nodes + n) {}
    const std::vector<int> &getVectorNodeList() const { return nodes_; }
  private:
    std::vector<int> nodes_;
};

class Elements {
  public:
    void reserve(std::size_t n) { elems_.reserve(n); }
    void addElement(const FiniteElement &fe) { elems_.push_back(fe); }
    int numberElements() const { return (int)elems_.size(); }
    const FiniteElement &getElement(int i) const { return elems_[(std::size_t)i]; }
  private:
    std::vector<FiniteElement> elems_;
};

template <class SC, class LO, class GO, class NO>
class Domain {
  public:
    typedef Map<LO, GO, NO> Map_Type;
    typedef Teuchos::RCP<const Map_Type> MapConstPtr_Type;
    Domain(int dim, const std::string &fe) : dim_(dim), FEType_(fe) {}
    int getDimension() const { return dim_; }
    std::string getFEType() const { return FEType_; }
    Teuchos::RCP<Elements> getElementsC() const { return elementsC_; }
    Teuchos::RCP<std::vector<std::vector<double> > > getPointsRepeated() const { return pointsRep_; }
    MapConstPtr_Type getMapRepeated() const { return mapRepeated_; }
    MapConstPtr_Type getElementMap() const { return elementMap_; }   // P0 spaces: numbering of the elements (Domain_decl.hpp)
    int dim_;
    std::string FEType_;
    Teuchos::RCP<Elements> elementsC_;
    Teuchos::RCP<std::vector<std::vector<double> > > pointsRep_;
    MapConstPtr_Type mapRepeated_, elementMap_;
};

template <class SC, class LO, class GO, class NO>
class MultiVector {
  public:
    MultiVector(const SC *data, std::size_t n) : data_(data, data + n) {}
    int getNumVectors() const { return 1; }
    Teuchos::ArrayRCP<const SC> getData(int) const { return Teuchos::ArrayRCP<const SC>(data_.data(), data_.size()); }
    Teuchos::ArrayRCP<SC> getDataNonConst(int) { return Teuchos::ArrayRCP<SC>(data_.data(), data_.size()); }
  private:
    std::vector<SC> data_;
};

// what a fill-complete matrix is to a standalone program: structure + values, as seated by FE_b200
template <class SC, class LO, class GO, class NO>
class Matrix {
  public:
    Matrix() : fillComplete_(false) {}
    std::shared_ptr<const void> structure;   // b200::CsrStructure<GO>
    std::shared_ptr<double> values;
    std::size_t nnz = 0;
    bool fillComplete_;
};
} // namespace FEDD

#include "FE_b200.hpp"

namespace FEDD {
namespace b200 {
// seat step of the standalone containers: the matrix takes (shared) ownership of the structure and of the value buffer
template <class SC, class LO, class GO, class NO>
void seat_csr(Teuchos::RCP<Matrix<SC, LO, GO, NO> > &A, LocalCsr<SC, LO, GO> &csr, Teuchos::RCP<const Map<LO, GO, NO> >,
              Teuchos::RCP<const Map<LO, GO, NO> >, bool callFillComplete)
{
    if (A.is_null()) A = Teuchos::rcp(new Matrix<SC, LO, GO, NO>());
    A->structure = csr.pattern;
    A->values = csr.values;
    A->nnz = csr.nnz;
    A->fillComplete_ = callFillComplete;
}
} // namespace b200
} // namespace FEDD
