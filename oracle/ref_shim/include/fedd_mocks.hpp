// Mock containers with the interface FEDDLib's hot-path assembly routines read (Domain, Elements,
// FiniteElement, Map, Matrix, MultiVector).  Matrix forwards insertGlobalValues / fillComplete to the
// oracle's C emulation (fedd_oracle.c) so both code paths share one accumulate step.
#pragma once
#include <string>
#include <vector>

#include "teuchos_mock.hpp"

extern "C" {
struct fo_matrix;
int fo_insert(fo_matrix *A, int64_t row, int32_t n, const int64_t *cols, const double *vals);
int64_t fo_fill_complete(fo_matrix *A);
}

namespace FEDD {

template <class LO, class GO, class NO>
class Map {
  public:
    explicit Map(const GO *gids = nullptr, std::size_t n = 0) : gids_(gids, gids + n) {}
    GO getGlobalElement(LO i) const { return gids_.at(i); }
    std::size_t getNodeNumElements() const { return gids_.size(); }
  private:
    std::vector<GO> gids_;
};

class FiniteElement {
  public:
    FiniteElement() {}
    explicit FiniteElement(const std::vector<int> &n) : nodes_(n) {}
    int getNode(int i) const { return nodes_.at(i); }
    const std::vector<int> &getVectorNodeList() const { return nodes_; }
  private:
    std::vector<int> nodes_;
};

class Elements {
  public:
    void addElement(const FiniteElement &fe) { elems_.push_back(fe); }
    int numberElements() const { return (int)elems_.size(); }
    const FiniteElement &getElement(int i) const { return elems_.at(i); }
  private:
    std::vector<FiniteElement> elems_;
};

template <class SC, class LO, class GO, class NO> class Mesh { };
template <class SC, class LO, class GO, class NO> class MeshUnstructured { };

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
    MapConstPtr_Type getElementMap() const { return elementMap_; }
    void initializeFEData() {}
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

template <class SC, class LO, class GO, class NO>
class Matrix {
  public:
    typedef Map<LO, GO, NO> Map_Type;
    typedef Teuchos::RCP<Map_Type> MapPtr_Type;
    typedef Teuchos::RCP<const Map_Type> MapConstPtr_Type;
    explicit Matrix(fo_matrix *sink) : sink_(sink), fillCompleteCalls_(0) {}
    void insertGlobalValues(GO row, const Teuchos::ArrayView<GO> &cols, const Teuchos::ArrayView<SC> &vals)
    {
        static_assert(sizeof(GO) == sizeof(int64_t), "GO must be 64 bit");
        if (fo_insert(sink_, (int64_t)row, (int32_t)cols.size(), (const int64_t *)cols.getRawPtr(), vals.getRawPtr()) != 0)
            throw std::runtime_error("insertGlobalValues: row not in the row map");
    }
    void fillComplete() { fo_fill_complete(sink_); fillCompleteCalls_++; }
    void fillComplete(MapConstPtr_Type, MapConstPtr_Type) { fo_fill_complete(sink_); fillCompleteCalls_++; }
    fo_matrix *sink_;
    int fillCompleteCalls_;
};

} // namespace FEDD
