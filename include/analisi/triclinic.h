// triclinic.h -- TriclinicLammpsCell: a general 3x3 cell -> LAMMPS lower-triangular cell + rotation.
//
// Same role and public surface as the reference's lib/include/triclinic.h:5-80, which leans on
// Eigen's Householder QR.  Eigen is not available to this repository, so the 3x3 Householder QR is
// written out here (textbook algorithm; the reflector convention -- beta = -sign(c0)*norm,
// essential part = tail/(c0-beta), tau = (beta-c0)/beta -- and the order of the floating-point
// operations follow what a plain unblocked Householder QR does, so the result agrees with the
// reference's to the last bits; tests compare against fixtures produced by the compiled reference).
//
// Input: 9 doubles, the matrix M in C (row-major) order whose COLUMNS are the cell vectors a, b, c.
// M = Q R with R upper triangular, signs fixed so that diag(R) >= 0.  Then in the rotated frame
// a = (R00,0,0), b = (R01,R11,0), c = (R02,R12,R22): the LAMMPS cell is lx=R00, ly=R11, lz=R22,
// xy=R01, xz=R02, yz=R12, and a vector v becomes Q^T v.
#ifndef ANALISI_B200_TRICLINIC_H
#define ANALISI_B200_TRICLINIC_H

#include <cmath>
#include <cstring>
#include <limits>

template <class T>
class TriclinicLammpsCell {
public:
    using MatrixT = T;

    explicit TriclinicLammpsCell(const T *cell_) : cell(cell_) {
        // A[r][c] = M(r,c)
        T A[3][3];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) A[r][c] = cell_[3 * r + c];
        const bool diag = A[0][1] == 0 && A[0][2] == 0 && A[1][0] == 0 && A[1][2] == 0 && A[2][0] == 0 && A[2][1] == 0;
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) q[r][c] = r == c ? T(1) : T(0);
        if (diag) {
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) R[r][c] = A[r][c];
            is_diagonal = true;
            return;
        }
        is_diagonal = false;
        // --- unblocked Householder QR of A, in place: R in the upper triangle, essential parts below ---
        T tau[3] = {0, 0, 0};
        for (int k = 0; k < 3; ++k) {
            const int tail = 2 - k;   // entries below the diagonal in column k
            T tail_sq = 0;
            if (tail == 2)
                tail_sq = A[1][k] * A[1][k] + A[2][k] * A[2][k];
            else if (tail == 1)
                tail_sq = A[2][k] * A[2][k];
            const T c0 = A[k][k];
            T beta;
            if (tail_sq <= std::numeric_limits<T>::min()) {
                tau[k] = 0;
                beta = c0;
                for (int r = k + 1; r < 3; ++r) A[r][k] = 0;
            } else {
                beta = std::sqrt(c0 * c0 + tail_sq);
                if (c0 >= 0) beta = -beta;
                for (int r = k + 1; r < 3; ++r) A[r][k] = A[r][k] / (c0 - beta);
                tau[k] = (beta - c0) / beta;
            }
            A[k][k] = beta;
            // apply H_k = I - tau v v^T (v = [1, essential]) to the columns right of k
            if (tail == 0) {
                // a 1-row block: scaled by (1 - tau); there are no columns right of k = 2 anyway
            } else if (tau[k] != 0) {
                for (int c = k + 1; c < 3; ++c) {
                    T tmp = 0;
                    if (tail == 2)
                        tmp = A[k + 1][k] * A[k + 1][c] + A[k + 2][k] * A[k + 2][c];
                    else
                        tmp = A[k + 1][k] * A[k + 1][c];
                    tmp += A[k][c];
                    A[k][c] -= tau[k] * tmp;
                    for (int r = k + 1; r < 3; ++r) A[r][c] -= tau[k] * A[r][k] * tmp;
                }
            }
        }
        // --- Q = H_0 H_1 H_2 applied to the identity, last reflector first ---
        for (int k = 2; k >= 0; --k) {
            const int tail = 2 - k;
            if (tail == 0) {
                q[2][2] *= T(1) - tau[2];
            } else if (tau[k] != 0) {
                for (int c = k; c < 3; ++c) {
                    T tmp = 0;
                    if (tail == 2)
                        tmp = A[k + 1][k] * q[k + 1][c] + A[k + 2][k] * q[k + 2][c];
                    else
                        tmp = A[k + 1][k] * q[k + 1][c];
                    tmp += q[k][c];
                    q[k][c] -= tau[k] * tmp;
                    for (int r = k + 1; r < 3; ++r) q[r][c] -= tau[k] * A[r][k] * tmp;
                }
            }
        }
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) R[r][c] = c >= r ? A[r][c] : T(0);
        // --- make the diagonal of R non-negative: flip row i of R and column i of Q (reference reflect(), :28-37) ---
        for (int i = 0; i < 3; ++i) {
            if (R[i][i] < 0) {
                for (int c = 0; c < 3; ++c) R[i][c] = R[i][c] * T(-1);
                for (int r = 0; r < 3; ++r) q[r][i] = q[r][i] * T(-1);
            }
        }
    }

    // internal box row: xlo,ylo,zlo,lx/2,ly/2,lz/2(,xy,xz,yz)
    void set_lammps_cell(T *cel, bool triclinic = true) const {
        cel[0] = 0;
        cel[1] = 0;
        cel[2] = 0;
        cel[3] = R[0][0] / 2;
        cel[4] = R[1][1] / 2;
        cel[5] = R[2][2] / 2;
        if (triclinic) {
            cel[6] = R[0][1];
            cel[7] = R[0][2];
            cel[8] = R[1][2];
        }
    }
    bool is_same_cell(const T *other) const { return std::memcmp(other, cell, 9 * sizeof(T)) == 0; }
    // v <- Q^T v  (row vector times Q)
    void rotate_vec(T *v) const {
        const T a = v[0], b = v[1], c = v[2];
        for (int j = 0; j < 3; ++j) v[j] = a * q[0][j] + b * q[1][j] + c * q[2][j];
    }
    // Q as 9 doubles, column-major like the reference's Eigen matrix (:71-73): out[r + 3c] = Q(r,c)
    void getQ(T *out) const {
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) out[r + 3 * c] = q[r][c];
    }
    bool isDiagonal() const { return is_diagonal; }
    const T *getCell() const { return cell; }

private:
    const T *cell;
    T q[3][3], R[3][3];
    bool is_diagonal;
};

#endif
