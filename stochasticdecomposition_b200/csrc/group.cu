// group.cu -- several GPUs driven by ONE host thread, for hosts that are a single process like the reference's
// ("one host thread drives all 8 GPUs").  Pure orchestration over the per-device contexts: the
// observation shards, the replicated dual-side tables and the peer-memory exchange are exactly those of the
// one-process-per-GPU path (sharding.py); only the plumbing differs (direct peer pointers instead of CUDA IPC).
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>

#include "sdgpu_internal.cuh"

struct sdgpu_group {
	std::vector<sdgpu_ctx *> m;
	int64_t totalObs = 0;
	std::vector<int32_t> scratchI;
};

int sd_peer_attach_local(sdgpu_ctx **ctxs, int n);      // nccl_glue.cu

extern "C" int sdgpu_group_size(sdgpu_group *g) { return g ? (int) g->m.size() : 0; }
extern "C" sdgpu_ctx *sdgpu_group_member(sdgpu_group *g, int i) { return (g && i >= 0 && i < (int) g->m.size()) ? g->m[i] : nullptr; }

extern "C" void sdgpu_group_destroy(sdgpu_group *g) {
	if (!g) return;
	for (sdgpu_ctx *c : g->m) sdgpu_destroy(c);
	delete g;
}

extern "C" int sdgpu_group_create(const sdgpu_problem *prob, const sdgpu_caps *caps, int nDevices, const int *devices, sdgpu_group **out) {
	if (!prob || !caps || !out || nDevices < 1) return sdgpu_fail("sdgpu_group_create: bad argument");
	if (prob->num.rvdOmCnt > 0) return sdgpu_fail("sdgpu_group_create: random costs (rvdOmCnt > 0) are not supported in sharded mode");
	if (nDevices > sdgpu_ctx::kMaxPeers) return sdgpu_fail("sdgpu_group_create: at most %d devices", sdgpu_ctx::kMaxPeers);
	*out = nullptr;
	sdgpu_group *g = new sdgpu_group();
	for (int i = 0; i < nDevices; i++) {
		sdgpu_ctx *c = nullptr;
		if (sdgpu_create(prob, caps, devices ? devices[i] : i, &c) != 0) { sdgpu_group_destroy(g); return SDGPU_ERR; }
		g->m.push_back(c);
	}
	if (nDevices > 1 && sd_peer_attach_local(g->m.data(), nDevices) != 0) { sdgpu_group_destroy(g); return SDGPU_ERR; }
	*out = g;
	return 0;
}

extern "C" int sdgpu_group_reset(sdgpu_group *g) {
	if (!g) return sdgpu_fail("null group");
	for (sdgpu_ctx *c : g->m) if (sdgpu_reset(c) != 0) return SDGPU_ERR;
	// the exchange sequence number is NOT reset: the flags of the previous replication stay in the exchange buffers, and a
	// restarted sequence could meet a stale flag of the same value (the in-kernel wait would then read a stale slot)
	g->totalObs = 0;
	return 0;
}

extern "C" int sdgpu_group_get_counts(sdgpu_group *g, sdgpu_counts *out) {
	if (!g || !out) return sdgpu_fail("null argument");
	if (sdgpu_get_counts(g->m[0], out) != 0) return SDGPU_ERR;
	out->omega = g->totalObs;
	return 0;
}

// calcOmega (stocUpdate.c:326-348) over the shards: the first match in GLOBAL order wins; a new observation goes to member total % G
extern "C" int sdgpu_group_calc_omega(sdgpu_group *g, const double *observ, double tol, int *newOmegaFlag) {
	if (!g || !observ) return sdgpu_fail("null argument");
	const int G = (int) g->m.size();
	int64_t first = -1;
	for (int r = 0; r < G; r++) {
		int loc = sdgpu_omega_find(g->m[r], observ, tol);
		if (loc <= SDGPU_ERR) return SDGPU_ERR;
		if (loc >= 0) { int64_t glob = (int64_t) loc * G + r; if (first < 0 || glob < first) first = glob; }
	}
	if (first >= 0) {
		if (sdgpu_omega_bump(g->m[first % G], (int) (first / G), 1) != 0) return SDGPU_ERR;
		if (newOmegaFlag) *newOmegaFlag = 0;
		return (int) first;
	}
	const int owner = (int) (g->totalObs % G);
	int loc = sdgpu_omega_append(g->m[owner], observ, 1);
	if (loc < 0) return SDGPU_ERR;
	if (sdgpu_calc_delta(g->m[owner], 1, loc) != 0) return SDGPU_ERR;          // stocUpdate.c:25
	if (newOmegaFlag) *newOmegaFlag = 1;
	return (int) g->totalObs++;
}

extern "C" int sdgpu_group_update_dual(sdgpu_group *g, const double *pi, double mubBar, int currentIter, double tol,
		int *lambdaIdx, int *newLambdaFlag, int *sigmaIdx, int *newSigmaFlag) {
	if (!g || !pi) return sdgpu_fail("null argument");
	int li0 = 0, nl0 = 0, si0 = 0, ns0 = 0;
	for (size_t r = 0; r < g->m.size(); r++) {              // replicated tables: every member makes the same call and must agree
		int li, nl, si, ns;
		if (sdgpu_update_dual(g->m[r], pi, mubBar, currentIter, tol, &li, &nl, &si, &ns) != 0) return SDGPU_ERR;
		if (r == 0) { li0 = li; nl0 = nl; si0 = si; ns0 = ns; }
		else if (li != li0 || nl != nl0 || si != si0 || ns != ns0) return sdgpu_fail("group_update_dual: members disagree (replicated tables out of step)");
	}
	if (lambdaIdx) *lambdaIdx = li0;
	if (newLambdaFlag) *newLambdaFlag = nl0;
	if (sigmaIdx) *sigmaIdx = si0;
	if (newSigmaFlag) *newSigmaFlag = ns0;
	return 0;
}

extern "C" int sdgpu_group_basis_find_or_append(sdgpu_group *g, int retainBasis, int ck, int feasFlag, int sigmaIdx, int *newBasisFlag) {
	if (!g) return sdgpu_fail("null group");
	int b0 = 0, nb0 = 0;
	const int32_t s = sigmaIdx;
	for (size_t r = 0; r < g->m.size(); r++) {
		sdgpu_ctx *c = g->m[r];
		int nb = 1, b;
		// without random costs obsFeasible is constant true (randCost.c:208), so the dedup of stocUpdate.c:101-113 does not depend on the
		// observation: a member that holds none yet replays it on its basis list directly
		if (c->omegaCnt > 0) b = sdgpu_basis_find_or_append(c, retainBasis, 0, ck, feasFlag, 0, &s, nullptr, &nb);
		else {
			b = -1;
			if (!retainBasis)
				for (int64_t i = 0; i < c->basisCnt; i++)
					if (c->basis[i].feas && c->basis[i].phiLen == 0 && c->basis[i].sigmaIdx[0] == s) { b = (int) i; nb = 0; c->basis[i].weight++; break; }
			if (b < 0) b = sdgpu_basis_append(c, ck, feasFlag, 0, &s, nullptr);
		}
		if (b < 0) return SDGPU_ERR;
		if (r == 0) { b0 = b; nb0 = nb; }
		else if (b != b0 || nb != nb0) return sdgpu_fail("group_basis_find_or_append: members disagree");
	}
	if (newBasisFlag) *newBasisFlag = nb0;
	return b0;
}

extern "C" int sdgpu_group_sd_cut(sdgpu_group *g, const double *Xvect, int numSamples, int pi_eval_flag, double lb, sdgpu_cut *cut) {
	if (!g || !Xvect || !cut || !cut->beta) return sdgpu_fail("null argument");
	const int G = (int) g->m.size();
	if (G == 1) return sdgpu_sd_cut(g->m[0], Xvect, numSamples, pi_eval_flag, lb, cut);
	// launch every member's cut first (asynchronous: the merge kernels meet in the peer exchange), only then wait for any of them
	const unsigned seq0 = g->m[0]->peerSeq;
	int launchRc = 0;
	for (int r = 0; r < G && launchRc == 0; r++)
		launchRc = sdgpu_sd_cut_partial(g->m[r], Xvect, numSamples, pi_eval_flag, lb);
	if (launchRc != 0) {
		// a member failed to launch: the members already launched wait for it inside their merge kernels.  Every member that has not
		// taken part in exchange seq0 + 1 does so now with an error marker, so that all sequences stay in step and every kernel ends.
		const std::string why = g_sdgpu_err;
		for (int r = 0; r < G; r++) if (g->m[r]->peerSeq == seq0) sd_peer_poison_cut(g->m[r]);
		for (int r = 0; r < G; r++) { cudaSetDevice(g->m[r]->device); cudaStreamSynchronize(g->m[r]->stream); }
		return sdgpu_fail("group_sd_cut: a member failed to launch its cut (%s); the exchange was completed with an error marker", why.c_str());
	}
	int rc = 0;
	const int n1 = g->m[0]->n1;
	std::vector<double> beta((size_t) n1 + 1);
	int64_t maxLocal = 0;
	for (int r = 0; r < G; r++) maxLocal = std::max<int64_t>(maxLocal, g->m[r]->omegaCnt);
	g->scratchI.resize((size_t) std::max<int64_t>(1, maxLocal));
	for (int r = 0; r < G; r++) {
		sdgpu_cut part;
		part.beta = r == 0 ? cut->beta : beta.data();
		part.iStar = cut->iStar ? g->scratchI.data() : nullptr;
		int st = sdgpu_sd_cut_finish(g->m[r], numSamples, &part);
		if (st != 0) { rc = st; continue; }
		if (r == 0) { cut->alpha = part.alpha; cut->cummOld = part.cummOld; cut->cummAll = part.cummAll; }
		if (cut->iStar)
			for (int64_t l = 0; l < g->m[r]->omegaCnt; l++) cut->iStar[l * G + r] = g->scratchI[l];
	}
	cut->omegaCnt = (int32_t) g->totalObs; cut->numSamples = numSamples;
	return rc;
}
